timeout 900 python -m pytest tests/test_inference_gpu.py tests/test_model_gpu.py tests/test_conv_gpu.py -x -q -m gpu 2>&1 | tail -2
python profiles/inference_latency.py 2>&1 | tail -1 | cut -c1-330
ARGUS_PDL=0 python profiles/inference_latency.py 2>&1 | tail -1 | cut -c1-330
for i in 1 2; do python bench.py --steps 20 --warmup 4 --no-cpu-baseline --no-inference --no-torch-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['final_loss'], d['clocks']['sm_mhz'])"; done
