#!/bin/bash
# Wait-only build + ARGUS_PDL=1 diverges in every trial (race3_waitonly.json). Bisect: which kind of stream operation
# between two kernels must break the programmatic chain (ARGUS_PDL_BREAK bits: 1 = event wait, 2 = event record,
# 4 = memset / memcpy), and does the weight-gradient side stream matter?
export RACE_CONFIGS="default:1,default:0"
T=${1:-8}
run() { tag=$1; shift; env "$@" ARGUS_PDL=1 timeout 300 python profiles/experiments/race_matrix.py gpurun_out/race4_$tag.json $T 4 > gpurun_out/race4_$tag.log 2>&1; echo "== $tag"; cut -c1-260 gpurun_out/race4_$tag.log; }
run break0 ARGUS_PDL_BREAK=0
run break1 ARGUS_PDL_BREAK=1
run break2 ARGUS_PDL_BREAK=2
run break4 ARGUS_PDL_BREAK=4
run break7 ARGUS_PDL_BREAK=7
run break7_nooverlap ARGUS_PDL_BREAK=7 ARGUS_WGRAD_OVERLAP=0
run nooverlap ARGUS_PDL_BREAK=0 ARGUS_WGRAD_OVERLAP=0
