for W in 4 8 12 16; do echo "W=$W"; PHB_WARPS_PER_CTA=$W python tools/profile_target.py 700 900 exmouth 2 2>&1 | tail -1; done
