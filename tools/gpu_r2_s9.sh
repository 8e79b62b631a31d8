#!/bin/bash
# round 2, sessions 9-11 (N = 4, 2, 1 GPUs; N = $1): BASELINE.json configs[2] Abu Dhabi + REFINE, configs[4] Pilbara + REFINE
# (the 1/2/4/8 sweep; 8 is session 6), the strong-scaling bench at N, and the one-process host entries (C ABI) over N devices
N=${1:-4}
mkdir -p gpurun_out
T0=$SECONDS
L=gpurun_out/r2s9_n$N.log
nvidia-smi --query-gpu=index,name --format=csv | tee $L
if [ "$N" = "1" ]; then TR="python"; else TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541"; fi
if [ "$N" != "1" ]; then
  echo "== abudhabi + REFINE N=$N t=$((SECONDS-T0))s" | tee -a $L
  timeout 400 $TR tests/manual/run_scene.py --config abudhabi --refine --check 24 2>&1 | tail -1 | tee -a $L
  echo "== bench N=$N t=$((SECONDS-T0))s" | tee -a $L
  timeout 400 $TR bench.py --gpus $N --steps 4 --warmup 3 > gpurun_out/r2s9_bench_n$N.json 2> gpurun_out/r2s9_bench_n$N.err
  cat gpurun_out/r2s9_bench_n$N.json | tee -a $L; tail -2 gpurun_out/r2s9_bench_n$N.err | tee -a $L
  echo "== phb_invert_rows, exmouth, one process over $N devices t=$((SECONDS-T0))s" | tee -a $L
  timeout 400 python tests/manual/rows_e2e.py --config exmouth --scene-planes 2>&1 | tail -1 | tee gpurun_out/r2s9_rows_e2e_n$N.json | tee -a $L
fi
echo "== pilbara + REFINE N=$N t=$((SECONDS-T0))s" | tee -a $L
timeout 900 $TR tests/manual/run_scene.py --config pilbara --refine --check ${PILBARA_CHECK:-24} 2>&1 | tail -1 | tee -a $L
echo "done t=$((SECONDS-T0))s" | tee -a $L
