#!/bin/bash
# round 2, session 14 (8 GPUs): the strong-scaling bench at N=8 with the final library (e2e with the ranks waiting on the host)
mkdir -p gpurun_out
T0=$SECONDS
L=gpurun_out/r2s14.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541"
echo "== bench N=8" | tee $L
timeout 400 $TR bench.py --gpus 8 --steps 4 --warmup 3 > gpurun_out/r2s14_bench_n8.json 2> gpurun_out/r2s14_bench_n8.err
cat gpurun_out/r2s14_bench_n8.json | tee -a $L; tail -3 gpurun_out/r2s14_bench_n8.err | tee -a $L
echo "done t=$((SECONDS-T0))s" | tee -a $L
