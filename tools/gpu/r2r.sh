# r2r: full GPU test suite, lanes 2 vs 3, full default bench line
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
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
LANEARG="--lanes 2" run OD_X=0
LANEARG="--lanes 3" run OD_X=0
LANEARG="--lanes 4" run OD_X=0
timeout 900 python bench.py --check > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err; tail -c 600 gpurun_out/r2r_bench.err; python -c "
import json; d=json.loads(open('gpurun_out/r2r_bench.json').read().strip().splitlines()[-1]); print({k:d[k] for k in ('value','ms_per_step','gpu_launches','check','clocks')}); print(d['e2e']['value'], d['cpu_baseline']); print(json.dumps(d['extra'])[:1500]); print(json.dumps(d['scaling_b64'])[:600])"
