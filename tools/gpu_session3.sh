#!/bin/bash
# one short GPU call: full GPU parity suite on the product library, then throughput and parity of the
# software-pipelined term-loop variants (built with photic_b200.build.build_variant; PHB_LIB selects the library)
mkdir -p gpurun_out
T0=$SECONDS
timeout 200 python -m pytest tests -m gpu -x -q > gpurun_out/s3_pytest_gpu.log 2>&1; echo "pytest rc=$? t=$((SECONDS-T0))s" >> gpurun_out/s3_pytest_gpu.log
tail -4 gpurun_out/s3_pytest_gpu.log
for spec in libphotic_b200.so:16 libphotic_b200_pipe.so:16 libphotic_b200_pipe12.so:12; do
  lib=${spec%%:*}; W=${spec##*:}
  echo "== $lib W=$W t=$((SECONDS-T0))s" | tee -a gpurun_out/s3_speed.log
  PHB_LIB=$PWD/photic_b200/csrc/$lib PHB_WARPS_PER_CTA=$W timeout 90 python tools/profile_target.py 700 900 exmouth 2 2>&1 | tail -1 | tee -a gpurun_out/s3_speed.log
done
echo "== pipe parity t=$((SECONDS-T0))s" | tee -a gpurun_out/s3_speed.log
PHB_LIB=$PWD/photic_b200/csrc/libphotic_b200_pipe.so timeout 120 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "golden_scenes or seeded_scenes or objective_known" 2>&1 | tail -2 | tee -a gpurun_out/s3_speed.log
echo "== pipe12 parity t=$((SECONDS-T0))s" | tee -a gpurun_out/s3_speed.log
PHB_LIB=$PWD/photic_b200/csrc/libphotic_b200_pipe12.so PHB_WARPS_PER_CTA=12 timeout 120 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "golden_scenes or objective_known" 2>&1 | tail -2 | tee -a gpurun_out/s3_speed.log
echo "done t=$((SECONDS-T0))s" | tee -a gpurun_out/s3_speed.log
