"""Developer experiment: does ROI ORDER matter (L2 reuse between overlapping ROIs)? Same ROI set, stand-alone launch
with L2 flushed, in index order vs host-sorted by (image, level, y centre)."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from bench import rois_log_uniform, roialign_algorithmic_bytes, synth_inputs, DEPTH
from objectdetection_b200.config import config
from objectdetection_b200.maskrcnn import pyramid_roi_align
from objectdetection_b200 import utils

dev = torch.device("cuda", 0)
conf = config()
B, N = 2, 1000
rs = np.random.RandomState(5)
fmaps = [torch.from_numpy(rs.standard_normal((B, s, s, DEPTH)).astype(np.float32)).to(dev) for s in (256, 128, 64, 32)]
rois_np = rois_log_uniform(1234, B, N)
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

for P in (7, 14):
    out = torch.empty((1, B * N, P, P, DEPTH), dtype=torch.float32, device=dev)
    rois = torch.from_numpy(rois_np).to(dev)
    _, lv = pyramid_roi_align(fmaps, rois, conf.IMAGE_SHAPE, [P, P], out=out, return_levels=True)
    lv = lv.cpu().numpy()
    ab = roialign_algorithmic_bytes(rois_np, lv, P, DEPTH)["total"]
    yc = (rois_np[..., 0] + rois_np[..., 2]) / 2
    variants = {"index order": rois_np}
    for name, key in (("level,y sorted", lambda b: np.lexsort((yc[b], lv[b]))),
                      ("level desc,y sorted", lambda b: np.lexsort((yc[b], -lv[b]))),
                      ("y sorted only", lambda b: np.argsort(yc[b]))):
        variants[name] = np.stack([rois_np[b][key(b)] for b in range(B)])
    for name, r in variants.items():
        m, mn = timed(torch.from_numpy(np.ascontiguousarray(r)).to(dev), P, out)
        print(f"P={P:2d} {name:22s} mean {m:.4f} ms min {mn:.4f} ms  frac(mean) {ab / m / 1e6 / 6543.1:.3f}", flush=True)
