#!/bin/bash
# `ncu --set full` of the fused forward block tail (conv3 + BN + identity + ReLU + bit mask in the epilogue,
# the default) at the bench configuration. 13 such launches per step (3 + 4 + 6 bottlenecks of layers 1-3);
# step 1 is skipped. Run through gpurun AFTER the plain command exited 0.
set -x
TAG=${1:-r2}
python profiles/one_step.py 3 > gpurun_out/plain_ft_${TAG}.log 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k regex:"conv_gemm_kernel<.int.256, .int.0, .int.2, .int.11>" -s 13 -c 13 -o gpurun_out/prof_ft_${TAG} python profiles/one_step.py 3 \
    > gpurun_out/ncu_ft_${TAG}.log 2>&1
ncu -i gpurun_out/prof_ft_${TAG}.ncu-rep --page raw --csv > gpurun_out/ncu_raw_ft_${TAG}.csv 2>/dev/null
ncu -i gpurun_out/prof_ft_${TAG}.ncu-rep --page source --csv --kernel-name-base demangled -k regex:"conv_gemm_kernel<.int.256, .int.0, .int.2, .int.11>" > gpurun_out/ncu_source_ft_${TAG}.csv 2>/dev/null
rm -f gpurun_out/prof_ft_${TAG}.ncu-rep
du -sh gpurun_out; tail -3 gpurun_out/ncu_ft_${TAG}.log
