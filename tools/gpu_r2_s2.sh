#!/bin/bash
# round 2, session 2: GPU suite on the restored tree (incl. the mined non-converged / restart / REFINE goldens), baseline bench line
mkdir -p gpurun_out
T0=$SECONDS
L=gpurun_out/r2s2.log
echo "== gpu suite" | tee $L
timeout 700 python -m pytest tests -q -m gpu -x 2>&1 | tail -5 | tee -a $L
echo "== bench t=$((SECONDS-T0))s" | tee -a $L
timeout 300 python bench.py --steps 4 --warmup 3 > gpurun_out/r2s2_bench_n1.json 2> gpurun_out/r2s2_bench_n1.err
cat gpurun_out/r2s2_bench_n1.json | tee -a $L
echo "done t=$((SECONDS-T0))s" | tee -a $L
