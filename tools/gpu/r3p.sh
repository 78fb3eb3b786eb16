# r3p: the driver's command line is --steps 20 --warmup 5: which lane count is best when fill/drain is 1/5 of the region?
for L in 2 3 4 5 6 8; do
  for rep in 1 2; do
    python3 bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-extras --lanes $L 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('lanes', $L, round(d['value']), round(d['ms_per_step'],4), 'frac', round(d['roofline']['frac'],3))"
  done
done
