# r3v: the all_gather no longer joins the lane every step: N=2 at the driver's --steps 20 and at 300 steps
for K in 20 20 300; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2952$((RANDOM%10)) bench.py --gpus 2 --steps $K --warmup 5 --check --no-extras 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N=2 K=$K', round(d['value']), round(d['ms_per_step'],4), d['check'], 'e2e', round(d['e2e']['value'],1))"
done
python3 bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N=1 K=20', round(d['value']), round(d['ms_per_step'],4))"
timeout 600 python -m pytest tests/test_gpu_dist_nccl.py -m gpu -q 2>&1 | tail -1
