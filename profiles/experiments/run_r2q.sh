F="--steps 4 --warmup 3 --no-cpu-baseline --no-inference --no-torch-baseline"
export ARGUS_BENCH_TRACE=2
for i in 1 2 3; do python bench.py $F 2>&1 >/dev/null | grep -E "LOSS_TRACE|SIG [3-7] " ; echo; done
