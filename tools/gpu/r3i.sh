# r3i: flat 7x7 kernel at higher occupancy (min CTAs per SM 10 / 12 instead of 1: 51 / 42 registers)
run() {
  echo "=== $*"
  env "$@" timeout 300 python bench.py --steps 300 --warmup 5 --no-cpu-baseline --no-extras --lanes 4 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; s=d['roialign_standalone']
        print('step_ms', round(d['ms_per_step'],4), 'img/s', round(d['value']), 'serial', round(d['extra']['ms_per_step_one_at_a_time'],4), 'p14_ms', round(r['ms_per_launch'],4), 'frac', round(r['frac'],3), 'sa7', round(s['p7']['ms'],4), round(s['p7']['frac'],3), 'sa14', round(s['p14']['ms'],4), round(s['p14']['frac'],3))
    else: print(l[:300])
"
}
run OD_X=0
run ODHEAD_LIB=$PWD/objectdetection_b200/libodhead_minb10.so
run ODHEAD_LIB=$PWD/objectdetection_b200/libodhead_minb12.so
run OD_X=0
