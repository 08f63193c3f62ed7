#!/bin/bash
# Standard profile capture of one round (run on the GPU box through gpurun; writes into gpurun_out/).
# 1. plain run (must exit 0 without ncu)  2. launch list  3. --set full captures of the dominant kernels
set -x
TAG=${1:-r1}
python profiles/one_step.py 4 > gpurun_out/plain_${TAG}.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -s 1000 -c 480 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python profiles/one_step.py 4 > gpurun_out/ncu_list_${TAG}.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_gemm_kernel -s 400 -c 6 \
    -o gpurun_out/prof_conv_${TAG} python profiles/one_step.py 4 > gpurun_out/ncu_conv_${TAG}.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wgrad_kernel -s 130 -c 4 \
    -o gpurun_out/prof_wgrad_${TAG} python profiles/one_step.py 4 > gpurun_out/ncu_wgrad_${TAG}.log 2>&1
# DRAM bytes of every tensor-core launch of one step (~190 per step with the fused block tail's Gram GEMMs: skip two
# steps, capture two; conv_traffic.py keeps one period of the kernel sequence) -> roofline.traffic
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -k regex:"conv_gemm_kernel|wgrad_kernel|wgrad_xpose_kernel" -s 400 -c 400 --csv --log-file gpurun_out/conv_dram_${TAG}.csv \
    python profiles/one_step.py 5 > gpurun_out/ncu_dram_${TAG}.log 2>&1
ARGUS_PROFILE_DETAIL=1 python profiles/profile_detail.py > gpurun_out/detail_${TAG}.log 2>&1
python profiles/hbm_mix.py > gpurun_out/hbm_mix_${TAG}.log 2>&1
python profiles/micro_bn.py > gpurun_out/micro_bn_${TAG}.log 2>&1
tail -3 gpurun_out/ncu_conv_${TAG}.log gpurun_out/ncu_wgrad_${TAG}.log
