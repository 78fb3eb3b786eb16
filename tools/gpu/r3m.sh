# r3m: 7x7 and 14x14 ROIAlign of one step launched concurrently (same ROI order: do they share L2 lines?)
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "static_round_robin or processing_order" 2>&1 | tail -1
run() {
  echo "=== $ARGS $*"
  env "$@" timeout 300 python bench.py --steps 300 --warmup 5 --no-cpu-baseline --no-extras $ARGS 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print('step_ms', round(d['ms_per_step'],4), 'img/s', round(d['value']), 'serial', round(d['extra']['ms_per_step_one_at_a_time'],4), 'p14_ms', round(r['ms_per_launch'],4), 'frac', round(r['frac'],3))
    else: print(l[:300])
"
}
ARGS="--lanes 4" run OD_X=0
ARGS="--lanes 4 --roi-concurrent" run OD_X=0
ARGS="--lanes 1" run OD_X=0
ARGS="--lanes 1 --roi-concurrent" run OD_X=0
ARGS="--lanes 2 --roi-concurrent" run OD_X=0
