timeout 600 python -m pytest tests/test_model_gpu.py tests/test_conv_gpu.py -x -q -m gpu 2>&1 | tail -2
F="--steps 6 --warmup 3 --no-cpu-baseline --no-inference --no-torch-baseline"
export ARGUS_BENCH_TRACE=1
echo "-- BNRED=1, wgrad overlap off"
for i in 1 2 3 4; do ARGUS_BN_REDUCE_FUSED=1 ARGUS_WGRAD_OVERLAP=0 python bench.py $F 2>&1 >/dev/null | grep LOSS_TRACE | cut -d" " -f 7-12; done
echo "-- BNRED=1, ring kernels off"
for i in 1 2 3; do ARGUS_BN_REDUCE_FUSED=1 ARGUS_BN_RING=0 python bench.py $F 2>&1 >/dev/null | grep LOSS_TRACE | cut -d" " -f 7-12; done
