# r3s: caller-supplied ROI order (computed once per step, shared by the 7x7 and 14x14 calls)
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_goldens.py -m gpu -x -q -k "roi or crop" 2>&1 | tail -2
ODHEAD_LIB=$PWD/objectdetection_b200/libodhead_dbg.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "roi or crop" 2>&1 | tail -1
run() {
  echo "=== $ARGS $*"
  env "$@" timeout 300 python bench.py --steps 300 --warmup 5 --no-cpu-baseline --no-extras $ARGS 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; s=d['roialign_standalone']
        print('step_ms', round(d['ms_per_step'],4), 'img/s', round(d['value']), 'serial', round(d['extra']['ms_per_step_one_at_a_time'],4), 'p14_ms', round(r['ms_per_launch'],4), 'frac', round(r['frac'],3), 'sa7', round(s['p7']['ms'],4), round(s['p7']['frac'],3), 'sa14', round(s['p14']['ms'],4), round(s['p14']['frac'],3))
    else: print(l[:300])
"
}
ARGS="--lanes 4" run OD_X=0
ARGS="--lanes 4 --no-shared-order" run OD_X=0
ARGS="--lanes 4" run OD_X=1
ARGS="--lanes 4 --check" run OD_X=1
