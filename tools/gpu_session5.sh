#!/bin/bash
# extreme-parameter KAT on the product library; full GPU suite + throughput of the cold-out variant
mkdir -p gpurun_out
T0=$SECONDS
L=gpurun_out/s5.log
echo "== default: extreme KAT" | tee $L
timeout 100 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "extreme or objective_known" 2>&1 | tail -15 | tee -a $L
echo "== cold: full GPU suite t=$((SECONDS-T0))s" | tee -a $L
PHB_LIB=$PWD/photic_b200/csrc/libphotic_b200_cold.so timeout 200 python -m pytest tests -q -m gpu 2>&1 | tail -15 | tee -a $L
for lib in libphotic_b200.so libphotic_b200_cold.so; do
  echo "== $lib speed t=$((SECONDS-T0))s" | tee -a $L
  PHB_LIB=$PWD/photic_b200/csrc/$lib timeout 90 python tools/profile_target.py 700 900 exmouth 2 2>&1 | tail -1 | tee -a $L
done
echo "done t=$((SECONDS-T0))s" | tee -a $L
