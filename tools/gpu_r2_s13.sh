#!/bin/bash
# round 2, session 13 (1 GPU): state check -- full GPU suite, smoke(), bench line with the final library
mkdir -p gpurun_out
T0=$SECONDS
L=gpurun_out/r2s13.log
echo "== gpu suite" | tee $L
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -6 | tee -a $L
echo "== smoke t=$((SECONDS-T0))s" | tee -a $L
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee -a $L
echo "== bench t=$((SECONDS-T0))s" | tee -a $L
timeout 500 python bench.py --steps 5 --warmup 3 > gpurun_out/r2s13_bench_n1.json 2> gpurun_out/r2s13_bench_n1.err
cat gpurun_out/r2s13_bench_n1.json | tee -a $L; tail -3 gpurun_out/r2s13_bench_n1.err | tee -a $L
echo "== reference arm t=$((SECONDS-T0))s" | tee -a $L
timeout 500 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r2s13_bench_reference.json 2> gpurun_out/r2s13_bench_reference.err
cat gpurun_out/r2s13_bench_reference.json | tee -a $L
echo "done t=$((SECONDS-T0))s" | tee -a $L
