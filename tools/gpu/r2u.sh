# r2u: IoU quad split tests; does a half-SM persistent ROIAlign (1 CTA/SM) overlap better with the other lanes?
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_goldens.py -m gpu -x -q -k "target or Target" 2>&1 | tail -2
python tools/prof_cases.py cfg3 2>&1 | grep "us/iter" | head -3
run() {
  echo "=== $LANEARG $*"
  env "$@" timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --no-extras $LANEARG 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; s=d['roialign_standalone']
        print('step_ms', round(d['ms_per_step'],4), 'img/s', round(d['value']), 'serial', round(d['extra']['ms_per_step_one_at_a_time'],4), 'p14_ms', round(r['ms_per_launch'],4), 'frac', round(r['frac'],3))
    else: print(l[:300])
"
}
for L in 2 4; do
LANEARG="--lanes $L" run OD_X=0
LANEARG="--lanes $L" run OD_ROI_CPS=1 OD_ROI_RING_KB=180
LANEARG="--lanes $L" run OD_ROI_CPS=1 OD_ROI_RING_KB=100
LANEARG="--lanes $L" run OD_ROI_KERNEL=flat
done
