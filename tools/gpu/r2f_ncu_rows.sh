export OD_ROI_CPS=2 OD_ROI_RING_KB=96 OD_ROI_XPT=2
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:crop_rows -s 3 -c 1 -o gpurun_out/r2f_rows python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_ncu.log 2>&1
tail -2 gpurun_out/r2f_ncu.log
