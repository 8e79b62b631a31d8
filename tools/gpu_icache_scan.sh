# instruction-cache behaviour of the solve kernel at 4 and 16 warps per SM
M=sm__icc_requests.sum,sm__icc_requests_lookup_hit.sum,sm__icc_requests_lookup_miss.sum,sm__icc_requests_lookup_miss_tag_miss.sum,gcc__cache_requests_type_instruction.sum,gcc__cache_requests_type_instruction_lookup_miss.sum,gcc__gcc2xbar_requests_type_instruction.sum,gcc__cache_requests_type_constant.sum,gcc__cache_requests_type_constant_lookup_miss.sum,smsp__inst_executed.sum,gpu__time_duration.sum,idc__requests.sum,idc__requests_lookup_miss.sum
for W in 4 16; do
PHB_WARPS_PER_CTA=$W ncu --metrics $M --clock-control none -k regex:solve_kernel -c 1 --csv --log-file gpurun_out/icache_w$W.csv python tools/profile_target.py 128 160 > /dev/null 2>&1
echo "== W=$W"; awk -F'","' 'NR>2{print $(NF-2), $NF}' gpurun_out/icache_w$W.csv | tr -d '"'
done
