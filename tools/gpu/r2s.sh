# r2s: split DetectionTargetLayer kernels; the -m gpu suite against the OD_DEBUG_BOUNDS build; cfg3 timing
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "--- OD_DEBUG_BOUNDS build (libodhead_dbg.so), full -m gpu suite"
ODHEAD_LIB=$PWD/objectdetection_b200/libodhead_dbg.so timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -5 | tee gpurun_out/r2s_dbg_suite.log
ODHEAD_LIB=$PWD/objectdetection_b200/libodhead_dbg.so OD_ROI_MIN_POOL=1 OD_ROI_ORDER=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "roi or crop" 2>&1 | tail -2 | tee -a gpurun_out/r2s_dbg_suite.log
ODHEAD_LIB=$PWD/objectdetection_b200/libodhead_dbg.so OD_ROI_TMA_STORE=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "roi or crop" 2>&1 | tail -2 | tee -a gpurun_out/r2s_dbg_suite.log
python tools/time_configs.py 2>&1 | tail -25
