# throughput vs warps per CTA; usage: gpu_wscan.sh lib "W list"
LIB=${1:-photic_b200/csrc/libphotic_b200.so}
for W in ${2:-4 8 12 16}; do echo "W=$W ($LIB)"; PHB_LIB=$PWD/$LIB PHB_WARPS_PER_CTA=$W python tools/profile_target.py 700 900 exmouth 2 2>&1 | tail -1; done
