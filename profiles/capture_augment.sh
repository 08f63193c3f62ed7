#!/bin/bash
# ncu --set full of the augmentation kernels only (arc rasteriser, plasma mask, fused augmentation), text pages kept.
set -x
TAG=${1:-r2}
python profiles/one_step.py 3 > gpurun_out/plain_aug_${TAG}.log 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"augment_kernel|plasma_mask_kernel|arc_paint_kernel" \
    -s 3 -c 3 -o gpurun_out/prof_aug_${TAG} python profiles/one_step.py 3 > gpurun_out/ncu_aug_${TAG}.log 2>&1
ncu -i gpurun_out/prof_aug_${TAG}.ncu-rep --page raw --csv > gpurun_out/ncu_raw_aug_${TAG}.csv 2>/dev/null
ncu -i gpurun_out/prof_aug_${TAG}.ncu-rep --page details > gpurun_out/ncu_details_aug_${TAG}.txt 2>/dev/null
ncu -i gpurun_out/prof_aug_${TAG}.ncu-rep --page source --csv -k regex:augment_kernel > gpurun_out/ncu_source_augment_${TAG}.csv 2>/dev/null
rm -f gpurun_out/prof_aug_${TAG}.ncu-rep
