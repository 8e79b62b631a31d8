# parity (scene tests) and throughput of library variants: gpu_variants.sh "lib:W lib:W ..."
for spec in "$@"; do
  lib=${spec%%:*}; W=${spec##*:}
  echo "== $lib W=$W"
  PHB_LIB=$PWD/$lib PHB_WARPS_PER_CTA=$W timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "scene_inversion or determinism" 2>&1 | tail -1
  PHB_LIB=$PWD/$lib PHB_WARPS_PER_CTA=$W timeout 120 python tools/profile_target.py 700 900 exmouth 2 2>&1 | tail -1
done
