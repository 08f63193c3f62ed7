#!/bin/bash
# `ncu --set full` of the memory- / latency-bound (non-GEMM) kernels of one training step at the bench configuration
# (B = 256 pairs, 256x256). Run on the GPU box through gpurun AFTER the plain command exited 0; writes gpurun_out/.
# `-s` skips the matching launches of step 1 (warm-up), `-c` takes the first ones of step 2.
set -x
TAG=${1:-r2}
python profiles/one_step.py 3 > gpurun_out/plain_ng_${TAG}.log 2>&1 || exit 1
NCU="timeout 900 ncu --set full --clock-control none --import-source on"
# one launch per step each: augmentation, plasma pre-pass, max-pool, optimizer, loss, grad norm, weight pack, avg-pool
$NCU -k regex:"augment_kernel|plasma_|maxpool_fwd_kernel|clip_adam_kernel|pose_loss_kernel|grad_sqnorm_kernel|pack_weights_kernel|avgpool_fwd_kernel|spaghetti_draw_kernel" \
    -s 9 -c 10 -o gpurun_out/prof_single_${TAG} python profiles/one_step.py 3 > gpurun_out/ncu_single_${TAG}.log 2>&1
# stem backward (2 launches / step), head (3 fwd + 6 bwd launches / step)
$NCU -k regex:"stem_pool_bn_bwd_kernel|linear_fwd_kernel|linear_bwd_w_kernel|linear_bwd_x_kernel" \
    -s 11 -c 11 -o gpurun_out/prof_stemhead_${TAG} python profiles/one_step.py 3 > gpurun_out/ncu_stemhead_${TAG}.log 2>&1
# batch-norm forward apply: 48 per step; the first six of step 2 are layer1.0 (bn1, bn2, bn3 + downsample) and layer1.1
$NCU -k regex:"bn_apply_ring_kernel" -s 48 -c 6 -o gpurun_out/prof_bnapply_${TAG} python profiles/one_step.py 3 \
    > gpurun_out/ncu_bnapply_${TAG}.log 2>&1
# batch-norm backward: the LAST eight launches of step 2 are layer1 (largest tensors); 31 + 31 per step
$NCU -k regex:"bn_bwd_reduce_ring_kernel|bn_bwd_apply_ring_kernel" -s 116 -c 8 -o gpurun_out/prof_bnbwd_${TAG} \
    python profiles/one_step.py 3 > gpurun_out/ncu_bnbwd_${TAG}.log 2>&1
# the small finalize / reduce kernels (latency-bound): a handful of each
$NCU -k regex:"bn_finalize_kernel|bn_bwd_finalize_kernel|wgrad_reduce_kernel|bn_alg_" -s 180 -c 12 \
    -o gpurun_out/prof_small_${TAG} python profiles/one_step.py 3 > gpurun_out/ncu_small_${TAG}.log 2>&1
# gpurun copies back at most 64 MiB: keep the raw-metric / details pages as text, drop the binary reports
for f in single stemhead bnapply bnbwd small; do
  ncu -i gpurun_out/prof_${f}_${TAG}.ncu-rep --page raw --csv > gpurun_out/ncu_raw_${f}_${TAG}.csv 2>/dev/null
  ncu -i gpurun_out/prof_${f}_${TAG}.ncu-rep --page details > gpurun_out/ncu_details_${f}_${TAG}.txt 2>/dev/null
done
ncu -i gpurun_out/prof_single_${TAG}.ncu-rep --page source --csv -k regex:augment_kernel > gpurun_out/ncu_source_augment_${TAG}.csv 2>/dev/null
rm -f gpurun_out/prof_*_${TAG}.ncu-rep
du -sh gpurun_out
