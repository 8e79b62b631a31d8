#!/bin/bash
# round 2, session 3: GPU suite on the two-class / work-sharing kernel, bench line at N=1
mkdir -p gpurun_out
T0=$SECONDS
L=gpurun_out/r2s3.log
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv | tee $L
echo "== gpu suite" | tee -a $L
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -25 | tee -a $L
echo "== bench t=$((SECONDS-T0))s" | tee -a $L
timeout 400 python bench.py --steps 3 --warmup 3 > gpurun_out/r2s3_bench_n1.json 2> gpurun_out/r2s3_bench_n1.err
cat gpurun_out/r2s3_bench_n1.json | tee -a $L
tail -5 gpurun_out/r2s3_bench_n1.err | tee -a $L
echo "done t=$((SECONDS-T0))s" | tee -a $L
