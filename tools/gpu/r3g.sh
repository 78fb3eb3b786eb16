# r3g: batch-64 block with one CUDA graph per image group; the 8- and 16-images-per-GPU shapes of N=8 / N=4 on one GPU
for gb in 64 16 8; do
OD_BENCH_GLOBAL_B=$gb timeout 600 python bench.py --steps 50 --warmup 3 --no-cpu-baseline --check 2>gpurun_out/r3g_$gb.err | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('GLOBAL_B', $gb, d['scaling_b64'])"
tail -2 gpurun_out/r3g_$gb.err | cut -c1-200
done
OD_BENCH_GLOBAL_B=8 timeout 600 python bench.py --steps 50 --warmup 3 --no-cpu-baseline --no-graph 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('GLOBAL_B 8 no graph', d['scaling_b64'])"
