#!/bin/bash
# Follow-up: the early-trigger build diverges on a non-default stream (2 of 3 runs). Does breaking the programmatic chain
# after every event wait / event record / memory operation (ARGUS_PDL_BREAK=7) remove the divergence? And is the caller's
# stream PRIORITY needed at all ("plain" = a non-default stream of normal priority)?
TRIG=argus_b200/libargus_b200_trig.so
for i in 1 2 3 4; do
  ARGUS_B200_LIB=$TRIG ARGUS_PDL=1 ARGUS_PDL_BREAK=7 python profiles/experiments/race_locate.py prio gpurun_out/race2b_trig_break7_prio$i.json 6
done
for m in 1 2 4; do
  ARGUS_B200_LIB=$TRIG ARGUS_PDL=1 ARGUS_PDL_BREAK=$m python profiles/experiments/race_locate.py prio gpurun_out/race2b_trig_break${m}_prio.json 6
  ARGUS_B200_LIB=$TRIG ARGUS_PDL=1 ARGUS_PDL_BREAK=$m python profiles/experiments/race_locate.py prio gpurun_out/race2b_trig_break${m}_prio_b.json 6
done
