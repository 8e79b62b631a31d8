#!/bin/bash
# round 2, session 7 (1 GPU): compact slabs + persisting-L2 window (on/off), full GPU suite, ncu captures (Exmouth- and
# Qatar-shaped windows), the GPU side of configs[0] (Murion), bench line
mkdir -p gpurun_out
T0=$SECONDS
L=gpurun_out/r2s7.log
echo "== gpu suite" | tee $L
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -6 | tee -a $L
for p in 1 0; do
  echo "== speed, PHB_L2_PERSIST=$p t=$((SECONDS-T0))s" | tee -a $L
  PHB_L2_PERSIST=$p timeout 120 python tools/profile_target.py 700 900 exmouth 2 2>&1 | tail -1 | tee -a $L
done
echo "== speed, v7 library as committed before this change t=$((SECONDS-T0))s" | tee -a $L
PHB_LIB=$PWD/photic_b200/csrc/libphotic_b200_v7.so timeout 120 python tools/profile_target.py 700 900 exmouth 2 2>&1 | tail -1 | tee -a $L
echo "== ncu full, exmouth window t=$((SECONDS-T0))s" | tee -a $L
timeout 400 ncu --set full --clock-control none --import-source on -k regex:solve_kernel -c 1 -o gpurun_out/r02_solve_v8 -f python tools/profile_target.py 160 200 > gpurun_out/r2s7_ncu.log 2>&1
tail -2 gpurun_out/r2s7_ncu.log | tee -a $L
echo "== ncu full, qatar window t=$((SECONDS-T0))s" | tee -a $L
timeout 400 ncu --set full --clock-control none --import-source on -k regex:solve_kernel -c 1 -o gpurun_out/r02_solve_v8_qatar -f python tools/profile_target.py 160 200 qatar > gpurun_out/r2s7_ncu_qatar.log 2>&1
tail -2 gpurun_out/r2s7_ncu_qatar.log | tee -a $L
echo "== murion, one GPU t=$((SECONDS-T0))s" | tee -a $L
timeout 300 python tests/manual/run_scene.py --config murion --check 64 2>&1 | tail -1 | tee -a $L
echo "== bench t=$((SECONDS-T0))s" | tee -a $L
timeout 500 python bench.py --steps 5 --warmup 3 > gpurun_out/r2s7_bench_n1.json 2> gpurun_out/r2s7_bench_n1.err
cat gpurun_out/r2s7_bench_n1.json | tee -a $L
tail -3 gpurun_out/r2s7_bench_n1.err | tee -a $L
echo "done t=$((SECONDS-T0))s" | tee -a $L
