# r3a: full suite (F-RCNN rows now compared bit for bit), default bench with the chunked batch-64 block
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 900 python bench.py --check > gpurun_out/r3a_bench.json 2> gpurun_out/r3a_bench.err; tail -c 400 gpurun_out/r3a_bench.err
python -c "
import json; d=json.loads(open('gpurun_out/r3a_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','check','lanes')}); print('e2e', d['e2e']['value'], 'cpu', d['cpu_baseline']['value'], d['cpu_baseline']['value_1thread'])
print('roofline', {k:d['roofline'][k] for k in ('frac','ms_per_launch','traffic','with_two_steps_in_flight')})
print('standalone', d['roialign_standalone'])
print('b64', d['scaling_b64'])
print('extra', {k:(v if not isinstance(v,dict) else {kk:vv for kk,vv in v.items() if kk!='what' and kk!='cpu_baseline'}) for k,v in d['extra'].items()})"
