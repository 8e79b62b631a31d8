#!/bin/bash
# round 2, session 6 (8 GPUs): BASELINE.json configs[4] Pilbara (+REFINE) and configs[3] Qatar as whole scenes with run-time
# work sharing, and the strong-scaling bench at N=8 (e2e through phb_invert_host_multi)
mkdir -p gpurun_out
T0=$SECONDS
L=gpurun_out/r2s6.log
nvidia-smi --query-gpu=index,name --format=csv | tee $L
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541"
echo "== pilbara N=8 t=$((SECONDS-T0))s" | tee -a $L
timeout 420 $TR tests/manual/run_scene.py --config pilbara --refine --check 24 2>&1 | tail -2 | tee -a $L
echo "== qatar N=8 t=$((SECONDS-T0))s" | tee -a $L
timeout 300 $TR tests/manual/run_scene.py --config qatar --check 24 2>&1 | tail -2 | tee -a $L
echo "== bench N=8 t=$((SECONDS-T0))s" | tee -a $L
timeout 300 $TR bench.py --gpus 8 --steps 4 --warmup 3 > gpurun_out/r2s6_bench_n8.json 2> gpurun_out/r2s6_bench_n8.err
cat gpurun_out/r2s6_bench_n8.json | tee -a $L; tail -3 gpurun_out/r2s6_bench_n8.err | tee -a $L
echo "done t=$((SECONDS-T0))s" | tee -a $L
