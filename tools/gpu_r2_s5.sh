#!/bin/bash
# round 2, session 5 (2 GPUs): work sharing across real devices -- in-process (phb_invert_host_multi, peer access) and
# across processes (CUDA IPC under torch.distributed); strong-scaling bench at N=2 with and without sharing; Exmouth scene
mkdir -p gpurun_out
T0=$SECONDS
L=gpurun_out/r2s5.log
nvidia-smi --query-gpu=index,name --format=csv | tee $L
nvidia-smi topo -m 2>&1 | head -8 | tee -a $L
echo "== multi-device tests" | tee -a $L
timeout 400 python -m pytest tests/test_multi_device.py tests/test_host_shim.py -q -m gpu -x 2>&1 | tail -6 | tee -a $L
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541"
echo "== bench N=2 shared t=$((SECONDS-T0))s" | tee -a $L
timeout 500 $TR bench.py --gpus 2 --steps 3 --warmup 2 > gpurun_out/r2s5_bench_n2.json 2> gpurun_out/r2s5_bench_n2.err
cat gpurun_out/r2s5_bench_n2.json | tee -a $L; tail -3 gpurun_out/r2s5_bench_n2.err | tee -a $L
echo "== bench N=2 no sharing t=$((SECONDS-T0))s" | tee -a $L
timeout 500 $TR bench.py --gpus 2 --steps 3 --warmup 2 --no-share --no-e2e > gpurun_out/r2s5_bench_n2_noshare.json 2> gpurun_out/r2s5_bench_n2_noshare.err
cat gpurun_out/r2s5_bench_n2_noshare.json | tee -a $L; tail -3 gpurun_out/r2s5_bench_n2_noshare.err | tee -a $L
echo "== exmouth scene N=2 t=$((SECONDS-T0))s" | tee -a $L
timeout 400 $TR tests/manual/run_scene.py --config exmouth --check 24 2>&1 | tail -2 | tee -a $L
echo "done t=$((SECONDS-T0))s" | tee -a $L
