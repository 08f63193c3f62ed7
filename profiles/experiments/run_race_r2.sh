#!/bin/bash
# Round-2 localisation of the stream-ordering hazard: per-parameter gradient checksums of a default-stream
# run against high-priority-stream runs, with the early-PDL-trigger build (which reproduced the divergence 10 of 10 times
# in round 1) and with the shipped build.
TRIG=argus_b200/libargus_b200_trig.so
ARGUS_B200_LIB=$TRIG ARGUS_PDL=1 python profiles/experiments/race_locate.py default gpurun_out/race2_trig_default.json 6
for i in 1 2 3; do
  ARGUS_B200_LIB=$TRIG ARGUS_PDL=1 python profiles/experiments/race_locate.py prio gpurun_out/race2_trig_prio$i.json 6
done
python profiles/experiments/race_locate.py default gpurun_out/race2_default.json 6
for i in 1 2 3; do
  python profiles/experiments/race_locate.py prio gpurun_out/race2_prio$i.json 6
done
