# r3b: the bench line at N=8 and N=4 (what the driver's scaling run will see)
nvidia-smi topo -m 2>/dev/null | head -14
for N in 8 4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 300 --warmup 5 > gpurun_out/r3b_bench_n$N.json 2> gpurun_out/r3b_bench_n$N.err
tail -3 gpurun_out/r3b_bench_n$N.err | cut -c1-300
python -c "
import json; d=json.loads(open('gpurun_out/r3b_bench_n$N.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','n_gpus','ms_per_step','gpu_launches','lanes')})
print('e2e', {k:d['e2e'][k] for k in ('value','ms_per_step','h2d_ceiling_gbs','frac_of_ceiling','placement')})
print('roofline', d['roofline']['frac'], d['roofline']['ms_per_launch'])
print('b64', d['scaling_b64'])"
done
