#!/bin/bash
# Rate of the residual rare divergence (default stream, no programmatic launches) and what it depends on.
T=${1:-40}
run() { tag=$1; shift; env "$@" ARGUS_PDL=0 timeout 500 python profiles/experiments/race_matrix.py gpurun_out/race11_$tag.json $T 4 64 > gpurun_out/race11_$tag.log 2>&1; echo "== $tag"; grep -E "divergent|distinct" gpurun_out/race11_$tag.log | cut -c1-200; }
export RACE_CONFIGS="default:1,default:0"
run base
run noov ARGUS_WGRAD_OVERLAP=0
run nofusedred ARGUS_BN_REDUCE_FUSED=0
run noaug_ring0 ARGUS_BN_RING=0
