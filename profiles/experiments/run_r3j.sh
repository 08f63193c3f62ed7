# training launches wait at the top again (pdl_begin(0) == pdl_prologue); eval chain keeps the deferred wait
F="--steps 20 --warmup 4 --no-cpu-baseline --no-inference --no-torch-baseline"
for i in 1 2 3 4 5 6 7 8; do
python bench.py $F 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('base', d['ms_per_step'], d['final_loss'], d['clocks']['sm_mhz'])"
done
timeout 600 python -m pytest tests/test_inference_gpu.py -x -q -m gpu 2>&1 | tail -2
python profiles/inference_latency.py 2>&1 | tail -1 | cut -c1-330
