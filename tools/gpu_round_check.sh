set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r01b_pytest_gpu.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r01b_bench_n1.json 2> gpurun_out/r01b_bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01b_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r01b_bench_under_ncu.log 2>&1
tail -3 gpurun_out/r01b_pytest_gpu.log; cat gpurun_out/r01b_bench_n1.json
