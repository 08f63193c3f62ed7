# one-shot Gram quadratic form, fc bias gradient on the side stream: same bits expected (9.74431324005127), ms/step
timeout 600 python -m pytest tests/test_bn_algebra_gpu.py tests/test_model_gpu.py tests/test_train_gpu.py -x -q -m gpu 2>&1 | tail -2
F="--steps 20 --warmup 4 --no-cpu-baseline --no-inference --no-torch-baseline"
for i in 1 2 3 4; do
python bench.py $F 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('oneshot', d['ms_per_step'], d['final_loss'], d['clocks']['sm_mhz'], d['kernel_families_ms']['bn_algebra'])"
done
