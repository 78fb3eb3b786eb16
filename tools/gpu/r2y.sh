B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --lanes 1"
$B > gpurun_out/r2y_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:crop_rows -s 3 -c 1 -o gpurun_out/r2y_crop_rows $B > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:crop_bins -s 3 -c 1 -o gpurun_out/r2y_crop_bins $B > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2y_launches_bench_steps3.csv $B > /dev/null 2>&1
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
