"""Per-kernel CUDA times (torch.profiler) of the non-headline cases: python tools/prof_cases.py [nms100k] [frcnn] (developer aid)."""
import os
import sys

import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from objectdetection_b200 import fasterrcnn  # noqa: E402
from objectdetection_b200.proposals import non_max_suppression  # noqa: E402


def cu(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def run(name, fn, iters=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(iters):
            fn()
        torch.cuda.synchronize()
    print(f"== {name} ({iters} iterations)")
    for e in sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:12]:
        print(f"{e.device_time_total / iters:10.1f} us/iter  x{e.count / iters:<5.1f} {e.key[:100]}")


def main():
    cases = sys.argv[1:] or ["nms100k", "frcnn"]
    rs = np.random.RandomState(0)
    if "nms100k" in cases:
        n = 100000
        s = np.exp(rs.uniform(np.log(8), np.log(256), n))
        cy, cx = rs.uniform(0, 4096, n), rs.uniform(0, 4096, n)
        bx = cu((np.stack([cy - s / 2, cx - s / 2, cy + s / 2, cx + s / 2], 1) / 4096).astype(np.float32))[None]
        sc = cu(rs.random_sample(n).astype(np.float32))[None]
        run("NMS 100k boxes", lambda: non_max_suppression(bx, sc, n, 0.5))
    if "props64" in cases:   # 64 scan CTAs: long enough for ncu's PC sampler (python tools/prof_cases.py props64 under ncu -k regex:nms_scan)
        import _synth
        from objectdetection_b200 import Proposals, config, utils
        conf = config()
        shapes = utils.get_resnet_stage_shapes(conf, conf.IMAGE_SHAPE)
        B = 64
        anchors = utils.gen_anchors(conf.IMAGE_SHAPE, B, conf.RPN_ANCHOR_SCALES, conf.RPN_ANCHOR_RATIOS, shapes, conf.RESNET_STRIDES, conf.RPN_ANCHOR_STRIDE)
        probs, bbox = _synth.rpn_outputs(rs, B, anchors.shape[1])
        p, bb = cu(probs), cu(bbox)
        run("Proposals B=64", lambda: Proposals(conf, B, p, bb, anchors), iters=2)
    if "propsK" in cases:   # same 16 chunks scanned, different row length (bytes per chunk): K = 6000 / 3072 / 1536
        import _synth
        from objectdetection_b200 import Proposals, config, utils
        for Kpre in (6000, 3072, 1536):
            conf = config()
            conf.PRE_NMS_ROIS_COUNT = Kpre
            shapes = utils.get_resnet_stage_shapes(conf, conf.IMAGE_SHAPE)
            anchors = utils.gen_anchors(conf.IMAGE_SHAPE, 2, conf.RPN_ANCHOR_SCALES, conf.RPN_ANCHOR_RATIOS, shapes, conf.RESNET_STRIDES, conf.RPN_ANCHOR_STRIDE)
            probs, bbox = _synth.rpn_outputs(np.random.RandomState(0), 2, anchors.shape[1])
            p, bb = cu(probs), cu(bbox)
            run(f"Proposals B=2 pre-NMS {Kpre}", lambda: Proposals(conf, 2, p, bb, anchors), iters=5)
    if "cfg3" in cases:      # BASELINE configs[2]: DetectionTargetLayer, 2000 proposals, 100 GT, 200 sampled ROIs, batch 8
        import _synth
        from objectdetection_b200 import config
        from objectdetection_b200.data_processor import BuildDetectionTargets
        conf = config()
        props, cls, gt, pp, pn = _synth.target_inputs(np.random.RandomState(3), 8, 2000, 100)
        t = [cu(x) for x in (props, cls, gt, pp, pn)]
        run("DetectionTargetLayer B=8 N=2000 G=100 R=200",
            lambda: BuildDetectionTargets(conf, t[0], t[1], t[2], perm_pos=t[3], perm_neg=t[4]), iters=5)
    if "frcnn" in cases:
        h, w, na = 38, 63, 9
        fp = cu(rs.random_sample((1, h, w, 2 * na)).astype(np.float32))
        fb = cu(rs.normal(0, 0.5, size=(1, h, w, 4 * na)).astype(np.float32))
        run("FasterRCNN proposals 12000 -> 2000", lambda: fasterrcnn.Proposals('train', fp, fb, image_shape=(600, 1000, 3), nms_threshold=0.7))


if __name__ == "__main__":
    main()
