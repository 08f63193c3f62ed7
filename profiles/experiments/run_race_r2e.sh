#!/bin/bash
# Controls (no programmatic edges at all) and the residual divergence with every non-kernel operation breaking the
# chain and no side stream: which phase's kernel -> kernel edges carry it?
T=${1:-8}
run() { tag=$1; shift; env "$@" timeout 300 python profiles/experiments/race_matrix.py gpurun_out/race5_$tag.json $T 4 > gpurun_out/race5_$tag.log 2>&1; echo "== $tag"; cut -c1-300 gpurun_out/race5_$tag.log; }
RACE_CONFIGS="default:1,default:0,prio:1" run control_nopdl ARGUS_PDL=0
export RACE_CONFIGS="default:0"
run b7_noov_fwd ARGUS_PDL=1 ARGUS_PDL_BREAK=7 ARGUS_WGRAD_OVERLAP=0 ARGUS_PDL_PHASE=1
run b7_noov_bwd ARGUS_PDL=1 ARGUS_PDL_BREAK=7 ARGUS_WGRAD_OVERLAP=0 ARGUS_PDL_PHASE=2
run b7_noov_other ARGUS_PDL=1 ARGUS_PDL_BREAK=7 ARGUS_WGRAD_OVERLAP=0 ARGUS_PDL_PHASE=4
run b7_noov_all ARGUS_PDL=1 ARGUS_PDL_BREAK=7 ARGUS_WGRAD_OVERLAP=0 ARGUS_PDL_PHASE=7
