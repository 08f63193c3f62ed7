F="--steps 20 --warmup 4 --no-cpu-baseline --no-inference --no-torch-baseline"
for i in 1 2; do
python bench.py $F 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('aug   ', d['ms_per_step'], d['final_loss'], d['clocks']['sm_mhz'])"
python bench.py $F --no-augmentation 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('no-aug', d['ms_per_step'], d['final_loss'], d['clocks']['sm_mhz'])"
done
timeout 900 python -m pytest tests/test_optin_modes_gpu.py tests/test_dp_gpu.py -x -q -m gpu 2>&1 | tail -3
