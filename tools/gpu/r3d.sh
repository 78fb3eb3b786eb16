# r3d: ROI order for the flat 7x7 kernel too (pre-pass latency hidden by the lanes?), more lanes
run() {
  echo "=== $LANEARG $*"
  env "$@" timeout 300 python bench.py --steps 300 --warmup 5 --no-cpu-baseline --no-extras $LANEARG 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; s=d['roialign_standalone']
        print('step_ms', round(d['ms_per_step'],4), 'img/s', round(d['value']), 'serial', round(d['extra']['ms_per_step_one_at_a_time'],4), 'p14_ms', round(r['ms_per_launch'],4), 'frac', round(r['frac'],3), 'sa7', round(s['p7']['ms'],4), round(s['p7']['frac'],3), 'sa14', round(s['p14']['ms'],4), round(s['p14']['frac'],3))
    else: print(l[:300])
"
}
LANEARG="--lanes 4" run OD_ROI_ORDER=1
LANEARG="--lanes 4" run OD_ROI_ORDER=2
LANEARG="--lanes 4" run OD_ROI_ORDER=1
LANEARG="--lanes 4" run OD_ROI_ORDER=2
LANEARG="--lanes 6" run OD_ROI_ORDER=1
LANEARG="--lanes 8" run OD_ROI_ORDER=1
LANEARG="--lanes 8" run OD_ROI_ORDER=2
