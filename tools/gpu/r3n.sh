# r3n: final record for the final library hash: ncu traffic + launch list + default bench + reference arm
bash tools/gpu/ncu_traffic.sh 2>&1 | tail -3
timeout 900 python bench.py --check > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; tail -c 300 gpurun_out/final_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_reference.json 2>/dev/null; cut -c1-400 gpurun_out/final_reference.json
