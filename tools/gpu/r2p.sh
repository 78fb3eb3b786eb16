# r2p: two steps in flight (lanes) x ROIAlign kernel choice
run() {
  echo "=== $*"
  env "$@" timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --no-extras $LANEARG 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; s=d['roialign_standalone']
        print('step_ms', round(d['ms_per_step'],4), 'img/s', round(d['value']), 'eager', round(d['extra']['eager_ms_per_step'],4), 'p14_pipeline_ms', round(r['ms_per_launch'],4), 'frac', round(r['frac'],3), 'sa7', round(s['p7']['ms'],4), 'sa14', round(s['p14']['ms'],4), round(s['p14']['frac'],3), 'e2e', round(d['e2e']['value'],1))
    else: print(l[:300])
"
}
export OD_ROI_TMA_STORE=0
LANEARG="--lanes 1" run OD_ROI_KERNEL=rows
LANEARG="--lanes 2" run OD_ROI_KERNEL=rows
LANEARG="--lanes 1" run OD_ROI_KERNEL=flat
LANEARG="--lanes 2" run OD_ROI_KERNEL=flat
LANEARG="--lanes 2 --check" run OD_ROI_KERNEL=flat
