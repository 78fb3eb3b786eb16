# r3l: the staged bulk-store variant with one CTA per SM (no register cap, no spills) - a fair comparison
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "roi or crop" 2>&1 | tail -1
OD_ROI_TMA_STORE=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "roi or crop" 2>&1 | tail -1
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
run OD_ROI_TMA_STORE=0
run OD_ROI_TMA_STORE=1
run OD_ROI_TMA_STORE=1 OD_ROI_RING_KB=190
export OD_ROI_TMA_STORE=1
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --lanes 1 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:crop_rows -s 3 -c 1 -o gpurun_out/r3l_crop_rows_tmast_cps1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --lanes 1 > /dev/null 2>&1
