# One `ncu --set full` capture of the 14x14 ROIAlign launch of the bench step + the bench line of the same build.
# Afterwards (in the container):  python tools/ncu_traffic.py gpurun_out/final_crop_rows.ncu-rep gpurun_out/final_plain.json
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --lanes 1"
$B > gpurun_out/final_plain.json 2> gpurun_out/final_plain.err || exit 1
ncu --set full --clock-control none --import-source on -k regex:crop_rows -s 3 -c 1 -o gpurun_out/final_crop_rows $B > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches_bench_steps3.csv $B > /dev/null 2>&1
ls -la gpurun_out/final_*
