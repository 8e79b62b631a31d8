for W in 4 16; do
PHB_WARPS_PER_CTA=$W ncu --section WarpStateStats --section SchedulerStats --section ComputeWorkloadAnalysis --clock-control none -k regex:solve_kernel -c 1 -o gpurun_out/stall_w$W -f python tools/profile_target.py 128 160 > gpurun_out/stall_w$W.log 2>&1
done
