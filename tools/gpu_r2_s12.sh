#!/bin/bash
# round 2, session 12 (2 GPUs): two CTAs of 8 warps per SM (with and without aligned evaluations) against the default on
# one GPU; then the 2-GPU leg of the sweeps: Abu Dhabi + REFINE, bench N=2, phb_invert_rows over 2 devices, Pilbara + REFINE
mkdir -p gpurun_out
T0=$SECONDS
L=gpurun_out/r2s12.log
echo "== speed: default | 2 CTAs x 8 warps | 2 x 8 aligned | 4 x 4 aligned  (exmouth, then qatar)" | tee $L
for cfg in "PHB_CTAS_PER_SM=1 PHB_ALIGN=0" "PHB_CTAS_PER_SM=2 PHB_ALIGN=0" "PHB_CTAS_PER_SM=2 PHB_ALIGN=1" "PHB_CTAS_PER_SM=4 PHB_ALIGN=1"; do
  echo "-- $cfg" | tee -a $L
  env $cfg timeout 120 python tools/profile_target.py 700 900 exmouth 2 2>&1 | tail -1 | tee -a $L
  env $cfg timeout 120 python tools/profile_target.py 700 900 qatar 2 2>&1 | tail -1 | tee -a $L
done
echo "== parity with 2 CTAs per SM, aligned t=$((SECONDS-T0))s" | tee -a $L
PHB_CTAS_PER_SM=2 PHB_ALIGN=1 timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "golden_scenes or seeded or mined or determinism" 2>&1 | tail -2 | tee -a $L
echo "== 2-GPU legs t=$((SECONDS-T0))s" | tee -a $L
PILBARA_CHECK=0 bash tools/gpu_r2_s9.sh 2
cat gpurun_out/r2s9_n2.log >> $L
echo "done t=$((SECONDS-T0))s" | tee -a $L
