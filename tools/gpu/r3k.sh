# r3k: why the staged bulk-store (TMA store) variant loses: one ncu --set full capture of it
export OD_ROI_TMA_STORE=1 OD_ROI_CPS=2
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --lanes 1"
$B > gpurun_out/r3k_plain.json 2>/dev/null || exit 1
ncu --set full --clock-control none --import-source on -k regex:crop_rows -s 3 -c 1 -o gpurun_out/r3k_crop_rows_tmast $B > /dev/null 2>&1
ls -la gpurun_out/r3k_*
