#!/bin/bash
# round 2, session 11 (1 GPU): L1 prefetch of the slab rows ahead of the centroid sum (default) against the library without
# it; parity of the new library; Pilbara + REFINE on ONE GPU with the library of the 4- and 8-GPU runs (the sweep's N = 1)
mkdir -p gpurun_out
T0=$SECONDS
L=gpurun_out/r2s11.log
echo "== speed: prefetch (default) / no prefetch; exmouth then qatar" | tee $L
for lib in libphotic_b200.so libphotic_b200_nopf.so; do
  PHB_LIB=$PWD/photic_b200/csrc/$lib timeout 120 python tools/profile_target.py 700 900 exmouth 2 2>&1 | tail -1 | tee -a $L
  PHB_LIB=$PWD/photic_b200/csrc/$lib timeout 120 python tools/profile_target.py 700 900 qatar 2 2>&1 | tail -1 | tee -a $L
done
echo "== parity t=$((SECONDS-T0))s" | tee -a $L
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "golden_scenes or seeded or mined or determinism or row_band" 2>&1 | tail -3 | tee -a $L
echo "== pilbara + REFINE N=1 (library without the prefetch = the kernel of the N = 4 and N = 8 runs) t=$((SECONDS-T0))s" | tee -a $L
PHB_LIB=$PWD/photic_b200/csrc/libphotic_b200_nopf.so timeout 1100 python tests/manual/run_scene.py --config pilbara --refine --check 0 2>&1 | tail -1 | tee -a $L
echo "done t=$((SECONDS-T0))s" | tee -a $L
