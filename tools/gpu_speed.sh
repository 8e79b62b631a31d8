# throughput only: gpu_speed.sh lib[:W] ...
for spec in "$@"; do
  lib=${spec%%:*}; W=16; [[ "$spec" == *:* ]] && W=${spec##*:}
  echo "== $lib W=$W"; PHB_LIB=$PWD/$lib PHB_WARPS_PER_CTA=$W timeout 120 python tools/profile_target.py 700 900 exmouth 2 2>&1 | tail -1
done
