# r2t: OD_DEBUG_BOUNDS suite; ncu evidence for decode / IoU-target kernels / final crop_rows; launch list of one step
ODHEAD_LIB=$PWD/objectdetection_b200/libodhead_dbg.so timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/r2t_dbg_suite.log
ODHEAD_LIB=$PWD/objectdetection_b200/libodhead_dbg.so OD_ROI_MIN_POOL=1 OD_ROI_ORDER=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "roi or crop" 2>&1 | tail -1 | tee -a gpurun_out/r2t_dbg_suite.log
ODHEAD_LIB=$PWD/objectdetection_b200/libodhead_dbg.so OD_ROI_TMA_STORE=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "roi or crop" 2>&1 | tail -1 | tee -a gpurun_out/r2t_dbg_suite.log
python tools/prof_cases.py cfg3 2>&1 | tail -8
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --lanes 1"
$B > gpurun_out/r2t_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2t_launches_bench_steps3.csv $B > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:crop_rows -s 3 -c 1 -o gpurun_out/r2t_crop_rows $B > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:proposal_decode -s 3 -c 1 -o gpurun_out/r2t_decode $B > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"detection_iou|detection_target" -s 4 -c 2 -o gpurun_out/r2t_targets python tools/prof_cases.py cfg3 > /dev/null 2>&1
ls -la gpurun_out/r2t_*
