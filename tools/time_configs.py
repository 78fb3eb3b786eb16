#!/usr/bin/env python
"""Times the C-ABI entry points on the BASELINE.json configurations that are not the bench line (configs 3-5) with
CUDA events (developer aid; prints a markdown table for profiles/). Run on a B200: python tools/time_configs.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _synth  # noqa: E402
from objectdetection_b200 import BuildDetectionTargets, DetectionLayer, Proposals, config, fasterrcnn, utils  # noqa: E402
from objectdetection_b200.data_processor import PreprareTrainData  # noqa: E402
from objectdetection_b200.maskrcnn import pyramid_roi_align  # noqa: E402
from objectdetection_b200.proposals import non_max_suppression  # noqa: E402


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def cu(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def main():
    conf = config()
    rs = np.random.RandomState(0)
    rows = []
    shapes = utils.get_resnet_stage_shapes(conf, conf.IMAGE_SHAPE)

    # config 3: training targets, batch 8, 2000 proposals, 100 GT, 200 ROIs (+ 28x28 mask targets)
    props, cls, gt, pp, pn = _synth.target_inputs(rs, 8, 2000, 100)
    masks = (rs.random_sample((8, 56, 56, 100)) > 0.5).astype(np.float32)
    a = [cu(x) for x in (props, cls, gt, pp, pn)]
    m = cu(masks)
    rows.append(("cfg3 DetectionTargetLayer B=8, 2000 proposals, 100 GT, R=200", timeit(lambda: BuildDetectionTargets(conf, a[0], a[1], a[2], perm_pos=a[3], perm_neg=a[4])), "8 images"))
    rows.append(("cfg3 + 28x28 mask targets", timeit(lambda: BuildDetectionTargets(conf, a[0], a[1], a[2], perm_pos=a[3], perm_neg=a[4], gt_masks=m)), "8 images"))

    # training proposals: batch 8, 6000 -> 2000
    for B in (2, 8, 64):
        anchors = utils.gen_anchors(conf.IMAGE_SHAPE, B, conf.RPN_ANCHOR_SCALES, conf.RPN_ANCHOR_RATIOS, shapes, conf.RESNET_STRIDES, conf.RPN_ANCHOR_STRIDE)
        probs, bbox = _synth.rpn_outputs(rs, B, anchors.shape[1])
        p, bb = cu(probs), cu(bbox)
        for training in (False, True):
            rows.append((f"Proposals B={B} 261888 anchors, 6000 -> {2000 if training else 1000}", timeit(lambda: Proposals(conf, B, p, bb, anchors, training=training)), f"{B} images"))
        if B == 64:
            P = Proposals(conf, B, p, bb, anchors).get_proposals()
            hp, hb = _synth.head_outputs(rs, B, 1000, 81)
            hp, hb = cu(hp), cu(hb)
            win = np.array([[131, 0, 893, 1024]] * B)
            rows.append((f"DetectionLayer B={B}, 1000 ROIs, 81 classes", timeit(lambda: DetectionLayer(conf, conf.IMAGE_SHAPE, B, win, P, hp, hb)), f"{B} images"))
            fm = [torch.rand((8, s, s, 256), device="cuda") for s in (256, 128, 64, 32)]
            out = torch.empty((1, 8 * 1000, 7, 7, 256), device="cuda")
            rows.append(("PyramidROIAlign 7x7 B=8 x 1000 ROIs", timeit(lambda: pyramid_roi_align(fm, P[:8], conf.IMAGE_SHAPE, [7, 7], out=out)), "8 images"))
            del fm, out

    # dense case: RPN-like outputs clustered around 40 objects (heavy overlap: the NMS has to visit every candidate and
    # keeps fewer than 1000) - the regime of a trained network, as opposed to the random boxes of the bench recipe
    B = 2
    anchors = utils.gen_anchors(conf.IMAGE_SHAPE, B, conf.RPN_ANCHOR_SCALES, conf.RPN_ANCHOR_RATIOS, shapes, conf.RESNET_STRIDES, conf.RPN_ANCHOR_STRIDE)
    anc = anchors[0].cpu().numpy()
    A = anc.shape[0]
    probs = np.zeros((B, A, 2), np.float32)
    bbox = np.zeros((B, A, 4), np.float32)
    acy, acx = (anc[:, 0] + anc[:, 2]) / 2, (anc[:, 1] + anc[:, 3]) / 2
    ah, aw = anc[:, 2] - anc[:, 0], anc[:, 3] - anc[:, 1]
    for b in range(B):
        objs = np.stack([rs.uniform(0.1, 0.9, 40), rs.uniform(0.1, 0.9, 40), rs.uniform(0.03, 0.3, 40), rs.uniform(0.03, 0.3, 40)], 1)
        d2 = ((acy[:, None] - objs[None, :, 0]) / objs[None, :, 2]) ** 2 + ((acx[:, None] - objs[None, :, 1]) / objs[None, :, 3]) ** 2
        j = d2.argmin(1)
        fg = np.exp(-d2.min(1) * 2) * np.exp(-np.abs(np.log(ah / objs[j, 2])) - np.abs(np.log(aw / objs[j, 3])))
        fg = np.clip(fg + rs.normal(0, 0.02, A), 0, 1).astype(np.float32)
        probs[b, :, 1], probs[b, :, 0] = fg, 1 - fg
        # regress towards the object (deltas / stddev), with noise
        bbox[b, :, 0] = (objs[j, 0] - acy) / ah / 0.1 + rs.normal(0, 0.3, A)
        bbox[b, :, 1] = (objs[j, 1] - acx) / aw / 0.1 + rs.normal(0, 0.3, A)
        bbox[b, :, 2] = np.log(objs[j, 2] / ah) / 0.2 + rs.normal(0, 0.3, A)
        bbox[b, :, 3] = np.log(objs[j, 3] / aw) / 0.2 + rs.normal(0, 0.3, A)
    p, bb = cu(probs), cu(np.clip(bbox, -20, 20).astype(np.float32))
    Pd = Proposals(conf, B, p, bb, anchors, DEBUG=True)
    kept = Pd.num_kept.cpu().numpy().tolist()
    rows.append((f"Proposals B=2, DENSE clustered boxes (NMS keeps {kept} of 6000, visits all)", timeit(lambda: Proposals(conf, B, p, bb, anchors)), "2 images"))

    # config 4: Faster R-CNN 600x1000, 12000 -> 2000, roi_pool 7x7 over 300 boxes, D=512
    h, w, na = 38, 63, 9
    fp = cu(rs.random_sample((1, h, w, 2 * na)).astype(np.float32))
    fb = cu(rs.normal(0, 0.5, size=(1, h, w, 4 * na)).astype(np.float32))
    rows.append(("cfg4 FasterRCNN Proposals 600x1000, 21546 anchors, 12000 -> 2000, thr 0.7", timeit(lambda: fasterrcnn.Proposals('train', fp, fb, image_shape=(600, 1000, 3), nms_threshold=0.7)), "1 image"))
    boxes = fasterrcnn.Proposals('train', fp, fb, image_shape=(600, 1000, 3), nms_threshold=0.7).get_proposals()[:300].contiguous()
    fmap = torch.rand((1, h, w, 512), device="cuda")
    rows.append(("cfg4 roi_pool 7x7 (crop 14x14 + max-pool), 300 boxes, D=512", timeit(lambda: fasterrcnn.roi_pool(fmap, boxes, (600, 1000, 3))), "300 ROIs"))

    # config 5: 100k-box NMS stress
    n = 100000
    s = np.exp(rs.uniform(np.log(8), np.log(256), n))
    cy, cx = rs.uniform(0, 4096, n), rs.uniform(0, 4096, n)
    bx = (np.stack([cy - s / 2, cx - s / 2, cy + s / 2, cx + s / 2], 1) / 4096).astype(np.float32)
    sc = rs.random_sample(n).astype(np.float32)
    bxc, scc = cu(bx)[None], cu(sc)[None]
    rows.append(("cfg5 NMS stress: 100,000 boxes, thr 0.5, max_out 100,000", timeit(lambda: non_max_suppression(bxc, scc, n, 0.5), iters=5, warm=1), "5.0e9 pairs"))

    # RPN targets, batch 8
    P8 = PreprareTrainData(conf)
    A = P8.anchors.shape[0]
    anc = P8.anchors.cpu().numpy()
    g8 = np.zeros((8, 100, 4))
    for b in range(8):
        g8[b] = np.round(np.clip(anc[rs.choice(A, 100, replace=False)] + rs.normal(0, 3, (100, 4)), 0, 1024))
        bad = (g8[b, :, 2] <= g8[b, :, 0]) | (g8[b, :, 3] <= g8[b, :, 1])
        g8[b, bad] = [100, 100, 164, 164]
    g8c = cu(g8)
    pp8 = torch.stack([torch.randperm(A, device="cuda") for _ in range(8)]).int()
    rows.append(("RPN targets B=8, 261888 anchors x 100 GT (fp64)", timeit(lambda: P8.build_rpn_targets(g8c, perm_pos=pp8, perm_neg=pp8), iters=10), "8 images"))

    print("| case | ms | per |")
    print("|---|---|---|")
    for name, ms, per in rows:
        print(f"| {name} | {ms:.3f} | {per} |")


if __name__ == "__main__":
    main()
