"""Regenerates tests/golden/reference_rpn_targets.npz by RUNNING THE REFERENCE'S OWN
`PreprareTrainData.build_rpn_targets` (MaskRCNN/building_blocks/data_processor.py:173-294, numpy float64).

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_rpn.py

The reference subsamples with the unseeded global `np.random.choice(idx, extra, replace=False)` (:254, :261). Here
the global legacy RNG is seeded before each call and the two permutations that `choice` draws internally
(`RandomState.permutation(len(idx))[:extra]`) are replayed from the same seed and stored, so the parity tests can
hand the very same permutations to the oracle / CUDA path (which take them as explicit inputs) and must reproduce
the reference's labels exactly, random subsampling included.
Cases: (a) the toy 128x128 configuration with the notebook's GT boxes (3 positives, first deltas
pinned), (b) a 256x256 case with 12 GT boxes where both subsampling branches fire (max_rpn_targets = 16).
"""
import contextlib
import io
import os
import sys
import tempfile
import types

import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_rpn_targets.npz")


def _import_reference():
    for name in ("tensorflow", "skimage", "skimage.transform", "keras", "keras.backend", "keras.layers"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["skimage"].transform = sys.modules["skimage.transform"]
    sys.modules["skimage.transform"].resize = lambda *a, **k: None
    sys.path.insert(0, REF)
    cwd = os.getcwd()
    os.chdir(tempfile.mkdtemp())  # the reference truncates ./logfile.log at import
    try:
        from MaskRCNN.building_blocks import data_processor as dp
        from MaskRCNN.building_blocks import utils as mutils
    finally:
        os.chdir(cwd)
    return dp, mutils


class _Conf:
    RESNET_STRIDES = [4, 8, 16, 32, 64]


def _run(dp, mutils, image, scales, gt, max_targets, seed):
    shapes = mutils.get_resnet_stage_shapes(_Conf, [image, image, 3])
    obj = object.__new__(dp.PreprareTrainData)        # __init__ only wires a dataset; set what the method reads
    obj.anchors = mutils.gen_anchors_pixel_coord(scales, [0.5, 1, 2], shapes, _Conf.RESNET_STRIDES, 1)
    obj.anchor_area = (obj.anchors[:, 2] - obj.anchors[:, 0]) * (obj.anchors[:, 3] - obj.anchors[:, 1])
    obj.max_rpn_targets = max_targets
    obj.bbox_std_dev = np.array([0.1, 0.1, 0.2, 0.2])
    np.random.seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        pos_anchors, cls, bbox = obj.build_rpn_targets(gt)
    # replay the permutations np.random.choice drew (legacy RandomState: permutation(pop_size)[:size])
    A = obj.anchors.shape[0]
    rs = np.random.RandomState(seed)
    # recompute the pre-subsampling label state exactly like the reference to know the population sizes
    gt_area = (gt[:, 2] - gt[:, 0]) * (gt[:, 3] - gt[:, 1])
    ov = np.stack([mutils.intersection_over_union(gt[i], obj.anchors, gt_area[i], obj.anchor_area) for i in range(len(gt))]).T
    amax = ov[np.arange(A), np.argmax(ov, 1)]
    lab = np.zeros(A, np.int32)
    lab[amax < 0.3] = -1
    lab[np.argmax(ov, 0)] = 1
    lab[amax >= 0.7] = 1
    n_pos0 = int((lab == 1).sum())
    perm_pos = np.arange(A, dtype=np.int32)
    extra = n_pos0 - max_targets // 2
    if extra > 0:
        p = rs.permutation(n_pos0)
        perm_pos[:n_pos0] = p
        lab[np.where(lab == 1)[0][p[:extra]]] = 0
    n_neg0 = int((lab == -1).sum())
    perm_neg = np.arange(A, dtype=np.int32)
    extra = n_neg0 - (max_targets - int((lab == 1).sum()))
    if extra > 0:
        p = rs.permutation(n_neg0)
        perm_neg[:n_neg0] = p
        lab[np.where(lab == -1)[0][p[:extra]]] = 0
    assert np.array_equal(lab, cls), "replayed permutations do not reproduce the reference's subsampling"
    return dict(anchors=obj.anchors, gt=gt.astype(np.float64), cls=cls.astype(np.int32), bbox=bbox, pos_anchors=pos_anchors,
                perm_pos=perm_pos, perm_neg=perm_neg, max_targets=np.array(max_targets), n_pos0=np.array(n_pos0),
                n_neg0=np.array(n_neg0))


def main():
    dp, mutils = _import_reference()
    g = {}
    # (a) toy config, notebook GT boxes (viz-iou-dummy.ipynb cell 15): G7 = 3 positives, 253 negatives, pinned deltas
    gt_a = np.array([[6, 73, 55, 124], [52, 46, 113, 107], [57, 30, 98, 71]], np.int32)
    a = _run(dp, mutils, 128, (8, 16, 32, 64, 128), gt_a, 256, seed=11)
    assert (a["cls"] == 1).sum() == 3 and (a["cls"] == -1).sum() == 253 and (a["cls"] == 0).sum() == 3836
    assert np.allclose(a["bbox"][0], [-0.78125, 0.78125, 1.23918082, 1.23918082], atol=1e-8)
    assert np.allclose(a["bbox"][1], [-0.234375, 0.390625, -1.33531393, -1.13528725], atol=1e-8)
    assert np.allclose(a["bbox"][2], [-1.49155337, 2.76213586, -1.97291405, 1.49282186], atol=1e-8)
    # (b) both subsampling branches: many GT boxes sitting exactly on anchors -> > max/2 positives
    rs = np.random.RandomState(3)
    shapes = mutils.get_resnet_stage_shapes(_Conf, [256, 256, 3])
    anc = mutils.gen_anchors_pixel_coord((16, 32, 64, 128, 256), [0.5, 1, 2], shapes, _Conf.RESNET_STRIDES, 1)
    pick = rs.choice(np.where((anc.min(1) >= 0) & (anc.max(1) <= 256))[0], 12, replace=False)
    gt_b = np.round(anc[pick]).astype(np.int32)
    b = _run(dp, mutils, 256, (16, 32, 64, 128, 256), gt_b, 16, seed=12)
    assert b["n_pos0"] > 8 and (b["cls"] == 1).sum() == 8 and (b["cls"] == -1).sum() == 8, (b["n_pos0"], (b["cls"] == 1).sum())
    for k, v in a.items():
        g["toy_" + k] = v
    for k, v in b.items():
        g["sub_" + k] = v
    np.savez_compressed(OUT, **g)
    print("wrote", OUT, {k: v.shape for k, v in g.items()})


if __name__ == "__main__":
    main()
