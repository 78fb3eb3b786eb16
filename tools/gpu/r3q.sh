# r3q: final: suites on both builds, smoke, ncu traffic of the final hash, driver-style bench
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -2 | tee gpurun_out/r3q_suite.log
ODHEAD_LIB=$PWD/objectdetection_b200/libodhead_dbg.so timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -2 | tee gpurun_out/r3q_dbg_suite.log
for v in "OD_ROI_TMA_STORE=1" "OD_ROI_TMA_STORE=1 OD_ROI_CPS=2" "OD_ROI_CPS=2" "OD_ROI_QPL=1" "OD_ROI_RING_KB=40" "OD_ROI_MIN_POOL=1" "OD_ROI_ORDER=0" "OD_ROI_ORDER=1" "OD_ROI_KERNEL=flat"; do
  echo "dbg build, $v: $(env $v ODHEAD_LIB=$PWD/objectdetection_b200/libodhead_dbg.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k 'roi or crop' 2>&1 | tail -1)" | tee -a gpurun_out/r3q_dbg_suite.log
done
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
bash tools/gpu/ncu_traffic.sh 2>&1 | tail -3
python3 bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r3q_bench20.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r3q_bench20.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['lib_source_hash'])"
