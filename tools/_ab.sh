for i in 1 2; do
for lib in objectdetection_b200/libodhead_prev.so objectdetection_b200/libodhead.so; do
  echo "== $lib"
  ODHEAD_LIB=$PWD/$lib python bench.py --steps 100 --warmup 5 --no-cpu-baseline --kernel-times 2>&1 | grep -E "us/step" | grep -E "nms_|total kernel|det_"
done
done
