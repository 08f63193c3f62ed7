timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/t_r2w.log 2>&1; echo "rc=$?" >> gpurun_out/t_r2w.log
tail -5 gpurun_out/t_r2w.log
python bench.py --steps 20 --warmup 5 > gpurun_out/b_r2w_1.json 2> gpurun_out/b_r2w_1.err; tail -c 600 gpurun_out/b_r2w_1.json
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-torch-baseline > gpurun_out/b_r2w_2.json 2> gpurun_out/b_r2w_2.err
python - <<'PY'
import json
for i in (1,2):
    d=json.loads(open(f"gpurun_out/b_r2w_{i}.json").read().strip().splitlines()[-1])
    print(i, d["ms_per_step"], d["value"], d["e2e"]["value"], d["final_loss"], d["clocks"], d["roofline_aggregate"]["frac"], d["hbm_step"])
PY
