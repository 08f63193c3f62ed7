#!/bin/bash
# After the epilogue fix (residual prefetch waits for the statistics pass of the previous chunk): every mode again.
T=${1:-12}
run() { tag=$1; shift; env "$@" timeout 400 python profiles/experiments/race_matrix.py gpurun_out/race7_$tag.json $T 4 64 > gpurun_out/race7_$tag.log 2>&1; echo "== $tag"; cut -c1-200 gpurun_out/race7_$tag.log; }
export RACE_CONFIGS="default:1,default:0,prio:1,plain:1"
run nopdl ARGUS_PDL=0
run nopdl_noov ARGUS_PDL=0 ARGUS_WGRAD_OVERLAP=0
run waitonly ARGUS_PDL=1
run trig ARGUS_PDL=1 ARGUS_B200_LIB=argus_b200/libargus_b200_trig.so
run trig_break7 ARGUS_PDL=1 ARGUS_PDL_BREAK=7 ARGUS_B200_LIB=argus_b200/libargus_b200_trig.so
