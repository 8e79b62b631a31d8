#!/bin/bash
# round 2, session 8 (1 GPU): full GPU suite, mapping study (warp per pixel vs teams of 2 / 4 / 8 warps) with ncu captures,
# ncu capture on a Qatar-shaped window, ncu launch list of the bench command
mkdir -p gpurun_out
T0=$SECONDS
L=gpurun_out/r2s8.log
echo "== gpu suite" | tee $L
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -8 | tee -a $L
echo "== mapping study t=$((SECONDS-T0))s" | tee -a $L
timeout 300 python tests/manual/mapping_study.py --reps 4000 2>&1 | tee gpurun_out/r2s8_mapping.jsonl | tee -a $L
echo "== speed t=$((SECONDS-T0))s" | tee -a $L
timeout 120 python tools/profile_target.py 700 900 exmouth 2 2>&1 | tail -2 | tee -a $L
echo "== ncu mapping study t=$((SECONDS-T0))s" | tee -a $L
timeout 600 ncu --set full --clock-control none --import-source on -k regex:eval_bench -o gpurun_out/r02_mapping -f python tests/manual/mapping_study.py --ncu-list --reps 600 > gpurun_out/r2s8_ncu_mapping.log 2>&1
tail -5 gpurun_out/r2s8_ncu_mapping.log | tee -a $L
echo "== ncu full, qatar window t=$((SECONDS-T0))s" | tee -a $L
timeout 400 ncu --set full --clock-control none --import-source on -k regex:solve_kernel -c 1 -o gpurun_out/r02_solve_v8_qatar -f python tools/profile_target.py 160 200 qatar > gpurun_out/r2s8_ncu_qatar.log 2>&1
tail -2 gpurun_out/r2s8_ncu_qatar.log | tee -a $L
echo "== ncu launch list of the bench command t=$((SECONDS-T0))s" | tee -a $L
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_ncu_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2s8_bench_under_ncu.log 2>&1
tail -3 gpurun_out/r02_ncu_launches_bench.csv | cut -c1-300 | tee -a $L
echo "done t=$((SECONDS-T0))s" | tee -a $L
