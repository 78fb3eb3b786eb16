# r3u: final code at N=4 (lanes, one gather stream, graph-captured batch-64 groups, shared ROI order)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 20 --warmup 5 --check > gpurun_out/r3u_bench_n4.json 2> gpurun_out/r3u_bench_n4.err
tail -2 gpurun_out/r3u_bench_n4.err | cut -c1-300
python -c "
import json; d=json.loads(open('gpurun_out/r3u_bench_n4.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','n_gpus','ms_per_step','gpu_launches','lanes','check')})
print('e2e', {k:d['e2e'][k] for k in ('value','ms_per_step','h2d_ceiling_gbs','frac_of_ceiling')})
print('roofline', d['roofline']['frac'], d['roofline']['ms_per_launch'], d['roofline']['traffic'])
print('b64', d['scaling_b64'])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29515 bench.py --impl reference --gpus 4 --steps 3 --warmup 1 2>/dev/null | tail -1 | cut -c1-300
