# r3h: record run - full suite (product build + OD_DEBUG_BOUNDS build), smoke, ncu evidence of the final build, default bench
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee gpurun_out/r3h_suite.log
ODHEAD_LIB=$PWD/objectdetection_b200/libodhead_dbg.so timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee gpurun_out/r3h_dbg_suite.log
for v in "OD_ROI_TMA_STORE=1 OD_ROI_CPS=2" "OD_ROI_CPS=2" "OD_ROI_QPL=1" "OD_ROI_RING_KB=40" "OD_ROI_MIN_POOL=1" "OD_ROI_ORDER=0" "OD_ROI_KERNEL=flat"; do
  echo "dbg build, $v: $(env $v ODHEAD_LIB=$PWD/objectdetection_b200/libodhead_dbg.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k 'roi or crop' 2>&1 | tail -1)" | tee -a gpurun_out/r3h_dbg_suite.log
done
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
bash tools/gpu/ncu_traffic.sh 2>&1 | tail -4
ncu --set full --clock-control none --import-source on -k regex:"detection_iou|detection_target" -s 4 -c 2 -o gpurun_out/final_targets python tools/prof_cases.py cfg3 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:crop_bins -s 3 -c 1 -o gpurun_out/final_crop_bins python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --lanes 1 > /dev/null 2>&1
python tools/prof_cases.py cfg3 2>&1 | grep "us/iter" | head -2
timeout 900 python bench.py --check --kernel-times > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; grep -A40 "launch order" gpurun_out/final_bench.err | head -60
