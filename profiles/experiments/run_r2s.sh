F="--steps 10 --warmup 3 --no-cpu-baseline --no-inference --no-torch-baseline"
for i in 1 2 3 4; do
ARGUS_FUSED_TAIL=1 ARGUS_BN_REDUCE_FUSED=0 python bench.py $F > gpurun_out/b_r2s_ft0_$i.json 2> gpurun_out/b_r2s_ft0_$i.err
ARGUS_FUSED_TAIL=1 ARGUS_BN_REDUCE_FUSED=1 python bench.py $F > gpurun_out/b_r2s_ft1_$i.json 2> gpurun_out/b_r2s_ft1_$i.err
done
python - <<'PY'
import json
for t in ["ft0","ft1"]:
  for i in (1,2,3,4):
    d=json.loads(open(f"gpurun_out/b_r2s_{t}_{i}.json").read().strip().splitlines()[-1])
    print(t, i, d["ms_per_step"], d["final_loss"], d["clocks"]["sm_mhz"])
PY
