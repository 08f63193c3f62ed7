#!/bin/bash
# The control of run_race_r2e.sh (no programmatic edges) also diverged: is it the weight-gradient side stream, the
# inspection kernels of the harness, or the batch size?
T=${1:-16}
run() { tag=$1; shift; env "$@" ARGUS_PDL=0 timeout 300 python profiles/experiments/race_matrix.py gpurun_out/race6_$tag.json $T 4 $B > gpurun_out/race6_$tag.log 2>&1; echo "== $tag"; cut -c1-300 gpurun_out/race6_$tag.log; }
export RACE_CONFIGS="default:0"
B=64 run full_ov1
B=64 run full_ov0 ARGUS_WGRAD_OVERLAP=0
B=64 run light_ov1 RACE_LIGHT=1
B=64 run light_ov0 RACE_LIGHT=1 ARGUS_WGRAD_OVERLAP=0
B=16 run light_ov1_b16 RACE_LIGHT=1
B=64 run full_ov0_noalg ARGUS_WGRAD_OVERLAP=0 ARGUS_BN_ALGEBRA=0
B=64 run full_ov0_nofusedred ARGUS_WGRAD_OVERLAP=0 ARGUS_BN_REDUCE_FUSED=0
