#!/bin/bash
# round check on one B200: GPU parity tests, the bench line, the ncu launch list of the same bench command
set -x
TAG=${1:-r01c}
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/${TAG}_pytest_gpu.log
python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'solve_kernel|classify_kernel|concat_queue|dfma_peak|refine' -c 200 --csv \
    --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_bench_under_ncu.log 2>&1
tail -3 gpurun_out/${TAG}_pytest_gpu.log; cat gpurun_out/${TAG}_bench_n1.json; cat gpurun_out/${TAG}_bench_reference.json
