# r2x: two GPUs - the NCCL gather test and the bench line at N=2 (lanes, one gather stream)
timeout 600 python -m pytest tests/test_gpu_dist_nccl.py -m gpu -q 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 300 --warmup 5 --check > gpurun_out/r2x_bench_n2.json 2> gpurun_out/r2x_bench_n2.err
tail -5 gpurun_out/r2x_bench_n2.err
python -c "
import json; d=json.loads(open('gpurun_out/r2x_bench_n2.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','n_gpus','ms_per_step','gpu_launches','check','collective','lanes')})
print('e2e', {k:d['e2e'][k] for k in ('value','ms_per_step','h2d_ceiling_gbs','frac_of_ceiling')})
print('roofline', d['roofline']['frac'], d['roofline']['ms_per_launch'])
print('b64', d['scaling_b64'])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 2>/dev/null | tail -1 | cut -c1-600
