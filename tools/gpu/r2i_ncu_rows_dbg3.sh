export OD_ROI_RING_KB=108 OD_ROI_TIMING_EXPERIMENT=3
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2i_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:crop_rows -s 3 -c 1 -o gpurun_out/r2i_rows python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2i_ncu.log 2>&1
tail -2 gpurun_out/r2i_ncu.log
