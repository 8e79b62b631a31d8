#!/bin/bash
# quick GPU check: parity tests + throughput of one mid-size inversion per library variant
# usage: gpu_quick.sh [lib1.so lib2.so ...]   (default: the product library)
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
libs="$@"; [ -z "$libs" ] && libs="photic_b200/csrc/libphotic_b200.so"
for lib in $libs; do
  echo "== $lib"
  PHB_LIB=$PWD/$lib python tools/profile_target.py 700 900 exmouth 2 2>&1 | tail -1
done
