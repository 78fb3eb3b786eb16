# r2z: ring-capacity fix (rows wider than half the ring take the flat path), small rings, y-loop unroll 2
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_goldens.py -m gpu -x -q -k "roi or crop" 2>&1 | tail -2
for kb in 32 40 64; do OD_ROI_RING_KB=$kb timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_goldens.py -m gpu -x -q -k "roi or crop" 2>&1 | tail -1; done
OD_ROI_RING_KB=40 OD_ROI_CPS=2 OD_ROI_TMA_STORE=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "roi or crop" 2>&1 | tail -1
run() {
  echo "=== $LANEARG $*"
  env "$@" timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --no-extras $LANEARG 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; s=d['roialign_standalone']
        print('step_ms', round(d['ms_per_step'],4), 'img/s', round(d['value']), 'serial', round(d['extra']['ms_per_step_one_at_a_time'],4), 'p14_ms', round(r['ms_per_launch'],4), 'frac', round(r['frac'],3), 'sa7', round(s['p7']['ms'],4), round(s['p7']['frac'],3), 'sa14', round(s['p14']['ms'],4), round(s['p14']['frac'],3))
    else: print(l[:300])
"
}
LANEARG="--lanes 4"
run OD_X=0
run ODHEAD_LIB=$PWD/objectdetection_b200/libodhead_yu2.so
run OD_X=0
run ODHEAD_LIB=$PWD/objectdetection_b200/libodhead_yu2.so
