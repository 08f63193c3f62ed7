#!/bin/bash
# Is the first trial of a process the outlier (uninitialised memory?) -- cluster sizes over all trials, two processes
# of the default build back to back, then the programmatic-launch builds.
T=${1:-10}
run() { tag=$1; shift; env "$@" timeout 400 python profiles/experiments/race_matrix.py gpurun_out/race8_$tag.json $T 4 64 > gpurun_out/race8_$tag.log 2>&1; echo "== $tag"; cut -c1-160 gpurun_out/race8_$tag.log; }
export RACE_CONFIGS="default:1,default:0,prio:1,prio:0"
run nopdl_a ARGUS_PDL=0
run nopdl_b ARGUS_PDL=0
run waitonly ARGUS_PDL=1
run trig ARGUS_PDL=1 ARGUS_B200_LIB=argus_b200/libargus_b200_trig.so
