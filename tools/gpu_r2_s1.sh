#!/bin/bash
# round 2, session 1: GPU suite on the carried-over build, mining of non-converged pixels, Murion on the GPU, baseline speed
mkdir -p gpurun_out
T0=$SECONDS
L=gpurun_out/r2s1.log
echo "== gpu suite" | tee $L
timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | tail -5 | tee -a $L
echo "== mining t=$((SECONDS-T0))s" | tee -a $L
timeout 900 python tests/manual/mine_nonconverged.py exmouth 0 3930 qatar 2000 3500 abudhabi 1000 2000 pilbara 6000 6600 2>&1 | tail -8 | tee -a $L
echo "== murion on one GPU t=$((SECONDS-T0))s" | tee -a $L
timeout 300 python tests/manual/run_scene.py --config murion --check 24 2>&1 | tail -1 | tee -a $L
echo "== speed t=$((SECONDS-T0))s" | tee -a $L
timeout 120 python tools/profile_target.py 700 900 exmouth 2 2>&1 | tail -1 | tee -a $L
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv | tee -a $L
echo "done t=$((SECONDS-T0))s" | tee -a $L
