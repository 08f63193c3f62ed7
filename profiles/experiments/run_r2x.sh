F="--steps 10 --warmup 3 --no-cpu-baseline --no-inference --no-torch-baseline"
run() { tag=$1; shift; for i in 1 2 3; do env "$@" python bench.py $F > gpurun_out/b_r2x_${tag}_$i.json 2> gpurun_out/b_r2x_${tag}_$i.err; done; }
run base X=1
run pdl ARGUS_PDL=1
run prio ARGUS_HIGH_PRIORITY_STREAM=1
run pdlprio ARGUS_PDL=1 ARGUS_HIGH_PRIORITY_STREAM=1
python - <<'PY'
import json
for t in ["base","pdl","prio","pdlprio"]:
  for i in (1,2,3):
    try:
      d=json.loads(open(f"gpurun_out/b_r2x_{t}_{i}.json").read().strip().splitlines()[-1])
      print(t, i, d["ms_per_step"], d["final_loss"], d["clocks"]["sm_mhz"])
    except Exception as e: print(t,i,"ERR",e)
PY
