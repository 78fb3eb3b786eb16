for v in 0 1 8 9 10 11 1 0; do
  ODHEAD_CROP_VARIANT=$v python bench.py --steps 60 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
s=d['roialign_standalone']
print('variant $v', 'step_ms %.4f'%d['ms_per_step'], 'p14pipe_ms %.4f frac %.3f'%(d['roofline']['ms_per_launch'], d['roofline']['frac']), 'sa_p7 %.4f (%.3f) sa_p14 %.4f (%.3f)'%(s['p7']['ms'],s['p7']['frac'],s['p14']['ms'],s['p14']['frac']))"
done
