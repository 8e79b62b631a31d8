#!/bin/bash
# round 2, sessions 16-17 (1 GPU): a kernel candidate (libphotic_b200.so) against the previous, validated library (_prev):
# full GPU suite on the candidate, speed of both on Exmouth- (6 dates) and Qatar-shaped (8 dates) rasters
mkdir -p gpurun_out
T0=$SECONDS
L=gpurun_out/${1:-r2s16}.log
echo "== gpu suite" | tee $L
timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -4 | tee -a $L
for lib in libphotic_b200.so libphotic_b200_prev.so; do
  echo "== speed $lib: exmouth, qatar t=$((SECONDS-T0))s" | tee -a $L
  PHB_LIB=$PWD/photic_b200/csrc/$lib timeout 100 python tools/profile_target.py 700 900 exmouth 2 2>&1 | tail -1 | tee -a $L
  PHB_LIB=$PWD/photic_b200/csrc/$lib timeout 100 python tools/profile_target.py 700 900 qatar 2 2>&1 | tail -1 | tee -a $L
done
echo "done t=$((SECONDS-T0))s" | tee -a $L
