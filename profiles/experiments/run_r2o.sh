F="--steps 10 --warmup 3 --no-cpu-baseline --no-inference --no-torch-baseline"
for i in 1 2 3; do
ARGUS_B200_LIB=argus_b200/libargus_b200_keepA.so python bench.py $F > gpurun_out/b_r2o_keepA_$i.json 2> gpurun_out/b_r2o_keepA_$i.err
ARGUS_BN_REDUCE_FUSED=2 python bench.py $F > gpurun_out/b_r2o_all_$i.json 2> gpurun_out/b_r2o_all_$i.err
done
python - <<'PY'
import json
for t in ["keepA","all"]:
  for i in (1,2,3):
    d=json.loads(open(f"gpurun_out/b_r2o_{t}_{i}.json").read().strip().splitlines()[-1])
    print(t, i, d["ms_per_step"], d["final_loss"])
PY
