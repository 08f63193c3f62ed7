#!/bin/bash
# Which ingredient makes early-triggering programmatic launches irreproducible? One process per library build.
T=${1:-10}
ARGUS_B200_LIB=argus_b200/libargus_b200_trig.so ARGUS_PDL=1 timeout 600 python profiles/experiments/race_matrix.py gpurun_out/race3_trig.json $T > gpurun_out/race3_trig.log 2>&1
ARGUS_B200_LIB=argus_b200/libargus_b200_trig_nonc.so ARGUS_PDL=1 timeout 600 python profiles/experiments/race_matrix.py gpurun_out/race3_trig_nonc.json $T > gpurun_out/race3_trig_nonc.log 2>&1
ARGUS_PDL=1 timeout 600 python profiles/experiments/race_matrix.py gpurun_out/race3_waitonly.json $T > gpurun_out/race3_waitonly.log 2>&1
tail -5 gpurun_out/race3_*.log
