F="--steps 10 --warmup 3 --no-cpu-baseline --no-inference --no-torch-baseline"
export ARGUS_BENCH_TRACE=1
for i in 1 2 3 4; do python bench.py $F 2>&1 >/dev/null | grep LOSS_TRACE; done
echo "-- no augmentation"
for i in 1 2 3; do python bench.py $F --no-augmentation 2>&1 >/dev/null | grep LOSS_TRACE; done
