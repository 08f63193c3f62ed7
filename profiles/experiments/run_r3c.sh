F="--steps 20 --warmup 4 --no-cpu-baseline --no-inference --no-torch-baseline"
for i in 1 2 3 4; do
python bench.py $F 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('base', d['ms_per_step'], d['final_loss'], d['clocks']['sm_mhz'])"
ARGUS_PDL=1 python bench.py $F 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('pdl ', d['ms_per_step'], d['final_loss'], d['clocks']['sm_mhz'])"
done
