#!/bin/bash
# round 2, session 10 (1 GPU): in-step vs out-of-phase evaluation loops (mapping_study --skews), CTA-wide alignment of the
# objective evaluations in the solve kernel (PHB_ALIGN=1) against the default, representative ncu capture (400 x 500 raster)
mkdir -p gpurun_out
T0=$SECONDS
L=gpurun_out/r2s10.log
echo "== mapping study with skews" | tee $L
timeout 300 python tests/manual/mapping_study.py --reps 3000 --skews 0,900,2777 2>&1 | tee gpurun_out/r2s10_mapping.jsonl | cut -c1-330 | tee -a $L
for a in 0 1; do
  echo "== speed, PHB_ALIGN=$a t=$((SECONDS-T0))s" | tee -a $L
  PHB_ALIGN=$a timeout 120 python tools/profile_target.py 700 900 exmouth 2 2>&1 | tail -1 | tee -a $L
  PHB_ALIGN=$a timeout 120 python tools/profile_target.py 700 900 qatar 2 2>&1 | tail -1 | tee -a $L
done
echo "== parity under PHB_ALIGN=1 t=$((SECONDS-T0))s" | tee -a $L
PHB_ALIGN=1 timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "golden_scenes or seeded or mined or depth_sigma or determinism" 2>&1 | tail -3 | tee -a $L
echo "== ncu full 400x500 default t=$((SECONDS-T0))s" | tee -a $L
timeout 600 ncu --set full --clock-control none --import-source on -k regex:solve_kernel -c 1 -o gpurun_out/r02_solve_v9 -f python tools/profile_target.py 400 500 > gpurun_out/r2s10_ncu.log 2>&1
tail -2 gpurun_out/r2s10_ncu.log | tee -a $L
echo "== ncu full 400x500 PHB_ALIGN=1 t=$((SECONDS-T0))s" | tee -a $L
PHB_ALIGN=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:solve_kernel -c 1 -o gpurun_out/r02_solve_v9_align -f python tools/profile_target.py 400 500 > gpurun_out/r2s10_ncu_align.log 2>&1
tail -2 gpurun_out/r2s10_ncu_align.log | tee -a $L
echo "done t=$((SECONDS-T0))s" | tee -a $L
