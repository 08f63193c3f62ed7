export ARGUS_PDL=1
python profiles/experiments/trajectory_repeat.py 1 1 16 8 256 2>&1 | grep -v LOSSES | tail -3
F="--steps 10 --warmup 3 --no-cpu-baseline --no-inference --no-torch-baseline"
for i in 1 2 3 4 5 6 7 8; do python bench.py $F 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['final_loss'])"; done
