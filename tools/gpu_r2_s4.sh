#!/bin/bash
# round 2, session 4: unified sand-only penalty (default) vs separate penalties, ncu capture of the solve kernel, bench line
mkdir -p gpurun_out
T0=$SECONDS
L=gpurun_out/r2s4.log
echo "== gpu suite (parity + multi device)" | tee $L
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_multi_device.py -q -m gpu -x 2>&1 | tail -6 | tee -a $L
for lib in libphotic_b200.so libphotic_b200_nounif.so; do
  echo "== $lib speed t=$((SECONDS-T0))s" | tee -a $L
  PHB_LIB=$PWD/photic_b200/csrc/$lib timeout 120 python tools/profile_target.py 700 900 exmouth 2 2>&1 | tail -1 | tee -a $L
done
echo "== ncu full t=$((SECONDS-T0))s" | tee -a $L
timeout 400 ncu --set full --clock-control none --import-source on -k regex:solve_kernel -c 1 -o gpurun_out/r02_solve_v7 -f python tools/profile_target.py 160 200 > gpurun_out/r2s4_ncu.log 2>&1
tail -2 gpurun_out/r2s4_ncu.log | tee -a $L
echo "== bench t=$((SECONDS-T0))s" | tee -a $L
timeout 500 python bench.py --steps 5 --warmup 3 > gpurun_out/r2s4_bench_n1.json 2> gpurun_out/r2s4_bench_n1.err
cat gpurun_out/r2s4_bench_n1.json | tee -a $L
tail -3 gpurun_out/r2s4_bench_n1.err | tee -a $L
echo "done t=$((SECONDS-T0))s" | tee -a $L
