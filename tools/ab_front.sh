#!/bin/bash
# A/B builds: live per-kernel timeline of the proposal front (developer aid, see tools/ab_libs.sh)
for name in "$@"; do
  lib=$PWD/objectdetection_b200/libodhead_$name.so
  [ "$name" = base ] && lib=$PWD/objectdetection_b200/libodhead.so
  echo "== $name"
  ODHEAD_LIB=$lib python bench.py --steps 50 --warmup 5 --no-cpu-baseline --kernel-times 2>&1 | grep -A12 "launch order" | grep -E "nms_|topk_" | cut -c1-80
done
