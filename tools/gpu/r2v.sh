# r2v: one persistent ROIAlign CTA per SM with a deep ring: ring size, quads per lane, 7x7 through the rows kernel
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_goldens.py -m gpu -x -q -k "roi or crop" 2>&1 | tail -2
OD_ROI_QPL=1 OD_ROI_MIN_POOL=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_goldens.py -m gpu -x -q -k "roi or crop" 2>&1 | tail -2
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
run OD_ROI_QPL=1
run OD_ROI_RING_KB=200
run OD_ROI_RING_KB=160
run OD_ROI_QPL=1 OD_ROI_RING_KB=180
run OD_ROI_MIN_POOL=1
run OD_ROI_MIN_POOL=1 OD_ROI_QPL=1
run OD_ROI_ORDER=1
run OD_ROI_L2_KEEP=2
LANEARG="--lanes 2" run OD_X=0
