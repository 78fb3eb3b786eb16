"""Developer experiment: does the ORDER's tail matter? Same ROI set walked (OD_ROI_ORDER=0, host-presorted) in
(a) level ascending, y   (what roi_order_kernel produces)   (b) expensive ROIs first, cheap (small P2) ROIs last."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from bench import rois_log_uniform, roialign_algorithmic_bytes, DEPTH
from objectdetection_b200.config import config
from objectdetection_b200.maskrcnn import pyramid_roi_align

dev = torch.device("cuda", 0)
conf = config()
B, N = 2, 1000
rs = np.random.RandomState(5)
fmaps = [torch.from_numpy(rs.standard_normal((B, s, s, DEPTH)).astype(np.float32)).to(dev) for s in (256, 128, 64, 32)]
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)

def timed(rois, P, out):
    ts = []
    for it in range(14):
        flush.fill_(float(it)); flush.sum()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); pyramid_roi_align(fmaps, rois, conf.IMAGE_SHAPE, [P, P], out=out); b.record()
        torch.cuda.synchronize()
        if it >= 2: ts.append(a.elapsed_time(b))
    return float(np.mean(ts)), float(np.min(ts))

recipes = {"SURVEY recipe 16..512 px": rois_log_uniform(1234, B, N),
           "proposal-like 8..160 px (mostly P2)": rois_log_uniform(99, B, N, lo=8, hi=160)}
for rname, rois_np in recipes.items():
    P = 14
    out = torch.empty((1, B * N, P, P, DEPTH), dtype=torch.float32, device=dev)
    _, lv = pyramid_roi_align(fmaps, torch.from_numpy(rois_np).to(dev), conf.IMAGE_SHAPE, [P, P], out=out, return_levels=True)
    lv = lv.cpu().numpy()
    ab = roialign_algorithmic_bytes(rois_np, lv, P, DEPTH)["total"]
    yc = (rois_np[..., 0] + rois_np[..., 2]) / 2
    size = np.sqrt((rois_np[..., 2] - rois_np[..., 0]) * (rois_np[..., 3] - rois_np[..., 1])) * 1024      # px
    band = np.minimum((yc * 32).astype(int), 31)
    small = (lv == 2) & (size < 40)
    grp_b = np.where(small, 99, -lv)                     # P5, P4, P3, P2-large first; small P2 last
    variants = {"index order": lambda b: np.arange(N),
                "level asc, y band": lambda b: np.lexsort((band[b], lv[b])),
                "expensive first, small P2 last, y band": lambda b: np.lexsort((band[b], grp_b[b])),
                "level desc, y band": lambda b: np.lexsort((band[b], -lv[b]))}
    print(f"--- {rname}: levels {np.bincount(lv.ravel(), minlength=6)[2:]}, small P2 {int(small.sum())}")
    for name, key in variants.items():
        r = np.stack([rois_np[b][key(b)] for b in range(B)])
        m, mn = timed(torch.from_numpy(np.ascontiguousarray(r)).to(dev), P, out)
        print(f"{name:42s} mean {m:.4f} ms min {mn:.4f} ms  frac(mean) {ab / m / 1e6 / 6543.1:.3f}", flush=True)
