# r2q: ROI processing order pre-pass (level, y band) + two lanes
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_goldens.py -m gpu -x -q 2>&1 | tail -3
OD_ROI_MIN_POOL=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_goldens.py -m gpu -x -q -k "roi or crop" 2>&1 | tail -2
run() {
  echo "=== $*"
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
export OD_ROI_TMA_STORE=0
LANEARG="--lanes 2" run OD_ROI_ORDER=0
LANEARG="--lanes 2" run OD_ROI_ORDER=1
LANEARG="--lanes 2" run OD_ROI_ORDER=1 OD_ROI_KERNEL=flat
LANEARG="--lanes 2 --check" run OD_ROI_ORDER=1 OD_ROI_MIN_POOL=1
LANEARG="--lanes 1" run OD_ROI_ORDER=1
