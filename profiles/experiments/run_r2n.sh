F="--steps 10 --warmup 3 --no-cpu-baseline --no-inference --no-torch-baseline"
for i in 1 2 3; do
ARGUS_BN_REDUCE_FUSED=0 python bench.py $F > gpurun_out/b_r2n_sep_$i.json 2> gpurun_out/b_r2n_sep_$i.err
python bench.py $F > gpurun_out/b_r2n_def_$i.json 2> gpurun_out/b_r2n_def_$i.err
ARGUS_FUSED_TAIL=1 python bench.py $F > gpurun_out/b_r2n_ft_$i.json 2> gpurun_out/b_r2n_ft_$i.err
done
python - <<'PY'
import json
for t in ["sep","def","ft"]:
  for i in (1,2,3):
    d=json.loads(open(f"gpurun_out/b_r2n_{t}_{i}.json").read().strip().splitlines()[-1])
    print(t, i, d["ms_per_step"], d["final_loss"])
PY
