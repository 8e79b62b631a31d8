#!/bin/bash
# N-GPU box: the in-process multi-device path of the C ABI on distinct devices (tests + one timed raster)
# usage: gpu_multi.sh rows cols [--multi-only]
mkdir -p gpurun_out
nvidia-smi -L | tee gpurun_out/s4_multi.log
if [ "$3" != "--multi-only" ]; then
  timeout 150 python -m pytest tests/test_multi_device.py tests/test_host_shim.py -x -q -m gpu 2>&1 | tail -3 | tee -a gpurun_out/s4_multi.log
fi
timeout 200 python tests/manual/multi_host.py ${1:-1400} ${2:-1000} $3 2>&1 | tail -2 | tee -a gpurun_out/s4_multi.log
