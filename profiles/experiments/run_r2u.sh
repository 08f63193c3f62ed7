F="--steps 6 --warmup 3 --no-cpu-baseline --no-inference --no-torch-baseline"
export ARGUS_BENCH_TRACE=1
for i in 1 2 3 4 5; do python bench.py $F 2>&1 >/dev/null | grep LOSS_TRACE | cut -d" " -f 7-12; done
