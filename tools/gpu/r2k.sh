export OD_ROI_RING_KB=108
python tools/gpu/r2k_sorted_rois.py
echo "--- flat kernel"
OD_ROI_KERNEL=flat python tools/gpu/r2k_sorted_rois.py
echo "--- gather_bw"
nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/gather_bw tools/ubench/gather_bw.cu && /tmp/gather_bw
