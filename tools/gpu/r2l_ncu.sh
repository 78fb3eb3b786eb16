# r2l: ncu --set full of crop_rows_kernel (new planner): full kernel, no-stores (exp 1), no-loads (exp 2), neither (exp 3)
export OD_ROI_RING_KB=108
for e in 0 1 2 3; do
  export OD_ROI_TIMING_EXPERIMENT=$e
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2l_plain$e.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:crop_rows -s 3 -c 1 -o gpurun_out/r2l_rows_e$e python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2l_ncu$e.log 2>&1
  tail -1 gpurun_out/r2l_ncu$e.log
done
