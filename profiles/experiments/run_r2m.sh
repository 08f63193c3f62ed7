set -x
timeout 600 python -m pytest tests/test_model_gpu.py tests/test_bn_algebra_gpu.py -x -q -m gpu > gpurun_out/t_r2m.log 2>&1; echo "rc=$?" >> gpurun_out/t_r2m.log
tail -4 gpurun_out/t_r2m.log
F="--steps 20 --warmup 4 --no-cpu-baseline --no-inference --no-torch-baseline"
for i in 1 2; do
python bench.py $F > gpurun_out/b_r2m_def_$i.json 2> gpurun_out/b_r2m_def_$i.err
ARGUS_FUSED_TAIL=1 python bench.py $F > gpurun_out/b_r2m_ft_$i.json 2> gpurun_out/b_r2m_ft_$i.err
done
python - <<'PY'
import json
for t in ["def_1","ft_1","def_2","ft_2"]:
    d=json.loads(open(f"gpurun_out/b_r2m_{t}.json").read().strip().splitlines()[-1])
    print(t, d["ms_per_step"], d["value"], d["e2e"]["value"], d["clocks"], d["final_loss"])
PY
