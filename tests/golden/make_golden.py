"""Regenerates tests/golden/reference_numpy.npz by RUNNING THE REFERENCE'S OWN numpy code.

Run in the build container only (needs /root/reference, which does not exist on the
GPU box):   python tests/golden/make_golden.py

The reference's TensorFlow layers cannot run here (TensorFlow is not installed), so
only its numpy-only functions are exercised:
  MaskRCNN/building_blocks/utils.py       gen_anchors, gen_anchors_pixel_coord, norm_boxes,
                                          denorm_boxes, get_resnet_stage_shapes,
                                          intersection_over_union, non_max_supression
  FasterRCNN/building_blocks/proposals.py get_anchors, corner_pixels_to_center_inv,
                                          FilterBoxes.clip_boxes/filter_min_size, non_max_suppression
  MaskRCNN/building_blocks/detection.py   unmold_detection
`tensorflow`, `skimage`, `keras`, `scipy` imports are stubbed (they are imported at module
top but not used by these functions).  The notebook constants G1–G8 of SURVEY.md §4 are
asserted while generating, so a drift of the reference is caught here.
"""
import contextlib
import hashlib
import io
import os
import sys
import tempfile
import types

import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_numpy.npz")


def _import_reference():
    for name in ("tensorflow", "skimage", "skimage.transform", "keras", "keras.backend", "keras.layers"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["skimage"].transform = sys.modules["skimage.transform"]
    if not hasattr(np, "int"):
        np.int = int  # FasterRCNN/building_blocks/proposals.py:137 uses the removed alias
    sys.path.insert(0, REF)
    cwd = os.getcwd()
    os.chdir(tempfile.mkdtemp())  # the reference truncates ./logfile.log at import
    try:
        from MaskRCNN.building_blocks import utils as mutils
        from FasterRCNN.building_blocks import proposals as fprops
        from MaskRCNN.building_blocks import detection as mdet
    finally:
        os.chdir(cwd)
    return mutils, fprops, mdet


class _Conf:
    RESNET_STRIDES = [4, 8, 16, 32, 64]


def main():
    mutils, fprops, mdet = _import_reference()
    g = {}
    ratios = [0.5, 1, 2]
    strides = [4, 8, 16, 32, 64]

    # ---- G8 / stage shapes
    shp128 = mutils.get_resnet_stage_shapes(_Conf, [128, 128, 3])
    shp1024 = mutils.get_resnet_stage_shapes(_Conf, [1024, 1024, 3])
    assert shp128.tolist() == [[32, 32], [16, 16], [8, 8], [4, 4], [2, 2]]
    g["stage_shapes_128"], g["stage_shapes_1024"] = shp128, shp1024

    # ---- G1: COCO anchors @1024^2 (4 MB -> keep a strided sample, a digest and the extrema)
    a1024 = mutils.gen_anchors([1024, 1024, 3], 1, (32, 64, 128, 256, 512), ratios, shp1024, strides, 1)
    assert a1024.shape == (1, 261888, 4) and a1024.dtype == np.float32
    assert abs(a1024.min() - (-0.353899)) < 1e-6 and abs(a1024.max() - 1.2913378) < 1e-6
    g["anchors_1024_shape"] = np.array(a1024.shape)
    g["anchors_1024_minmax"] = np.array([a1024.min(), a1024.max()], np.float32)
    g["anchors_1024_sample_rows"] = np.arange(0, 261888, 97)
    g["anchors_1024_sample"] = a1024[0, ::97].copy()
    g["anchors_1024_sha256"] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(a1024[0]).tobytes()).digest(), np.uint8)

    # ---- G2/G3: toy anchors @128^2, pixel coordinates, float64
    toy_scales = (8, 16, 32, 64, 128)
    apix = mutils.gen_anchors_pixel_coord(toy_scales, ratios, shp128, strides, 1)
    assert apix.shape == (4092, 4)
    assert abs(apix.min() - (-90.5096679919)) < 1e-9 and abs(apix.max() - 154.509667992) < 1e-9
    assert np.allclose(apix[[3970, 4054, 4074]],
                       [[64, 32, 96, 64], [0, 64, 64, 128], [50.745166, 41.372583, 141.254834, 86.627417]], atol=1e-6)
    g["anchors_toy_pixel"] = apix
    g["anchors_toy_norm"] = mutils.gen_anchors([128, 128, 3], 2, toy_scales, ratios, shp128, strides, 1)
    # anchor_stride 2 variant + non-square image
    shp_rect = mutils.get_resnet_stage_shapes(_Conf, [192, 320, 3])
    g["stage_shapes_192x320"] = shp_rect
    g["anchors_rect_s2_pixel"] = mutils.gen_anchors_pixel_coord((16, 32, 64, 128, 256), [0.5, 1, 2, 3], shp_rect, strides, 2)
    g["anchors_rect_s2_norm"] = mutils.gen_anchors([192, 320, 3], 1, (16, 32, 64, 128, 256), [0.5, 1, 2, 3], shp_rect, strides, 2)

    # ---- G4/G5: norm_boxes
    gt_px = np.array([[6, 73, 55, 124], [52, 46, 113, 107], [57, 30, 98, 71]])
    gt_n = mutils.norm_boxes(gt_px, (128, 128))
    assert np.allclose(gt_n[0], [0.04724409, 0.57480317, 0.42519686, 0.96850395], atol=1e-7)
    win = mutils.norm_boxes(np.array([131, 0, 893, 1024]), (1024, 1024))
    assert np.allclose(win, [0.12805474, 0., 0.87194526, 1.], atol=1e-7)
    g["norm_in_gt"], g["norm_out_gt"] = gt_px, gt_n
    g["norm_in_window"], g["norm_out_window"] = np.array([131, 0, 893, 1024]), win
    rs = np.random.RandomState(11)
    rnd_px = rs.randint(0, 1024, size=(64, 4))
    g["norm_in_rand"], g["norm_out_rand"] = rnd_px, mutils.norm_boxes(rnd_px, (800, 1024))
    rnd_n = rs.random_sample((64, 4))
    with contextlib.redirect_stdout(io.StringIO()):
        g["denorm_in_rand"], g["denorm_out_rand"] = rnd_n, mutils.denorm_boxes(rnd_n, (800, 1024))

    # ---- G6: numpy NMS of the shapes dataset + random cases
    sb = np.array([[26, 40, 86, 100], [6, 73, 56, 123], [52, 46, 112, 106], [57, 30, 97, 70]])
    keep = mutils.non_max_supression(sb, np.arange(4), 0.3)
    assert keep.tolist() == [3, 2, 1]
    g["npnms_boxes_0"], g["npnms_scores_0"], g["npnms_thr_0"], g["npnms_keep_0"] = sb, np.arange(4), np.array(0.3), keep
    for case in range(1, 6):
        n = 40 * case
        yx = rs.random_sample((n, 2)) * 100
        hw = rs.random_sample((n, 2)) * 40 + 2
        boxes = np.concatenate([yx, yx + hw], axis=1)
        scores = rs.permutation(n).astype(np.float64) / n      # distinct -> argsort order is unambiguous
        thr = [0.3, 0.5, 0.7, 0.1, 0.45][case - 1]
        g[f"npnms_boxes_{case}"], g[f"npnms_scores_{case}"] = boxes, scores
        g[f"npnms_thr_{case}"], g[f"npnms_keep_{case}"] = np.array(thr), mutils.non_max_supression(boxes, scores, thr)
    box = np.array([10., 10., 50., 60.])
    others = np.concatenate([rs.random_sample((32, 2)) * 40, rs.random_sample((32, 2)) * 40 + 40], axis=1)
    area = lambda b: (b[..., 2] - b[..., 0]) * (b[..., 3] - b[..., 1])
    g["npiou_box"], g["npiou_boxes"] = box, others
    g["npiou_out"] = mutils.intersection_over_union(box, others, area(box), area(others))

    # ---- unmold_detection (detection.py:8-53; numpy only)
    det = np.zeros((100, 6), np.float32)
    nd = 37
    yx = rs.random_sample((nd, 2)) * 0.6 + 0.13
    det[:nd, :4] = np.concatenate([yx, yx + rs.random_sample((nd, 2)) * 0.25], axis=1)
    det[5, 2], det[9, 3] = det[5, 0] - 0.002, det[9, 1] - 0.002     # collapse to zero area after rounding
    det[:nd, 4] = rs.randint(1, 81, nd)
    det[:nd, 5] = np.sort(rs.random_sample(nd))[::-1]
    with contextlib.redirect_stdout(io.StringIO()):
        ub, uc, us = mdet.unmold_detection((600, 800, 3), (1024, 1024, 3), det, np.array([131, 0, 893, 1024]))
    assert ub.shape[0] < nd
    g["unmold_in"], g["unmold_boxes"], g["unmold_class_ids"], g["unmold_scores"] = det, ub, uc, us

    # ---- Faster R-CNN numpy proposal layer pieces
    g["frcnn_base_anchors"] = fprops.get_anchors()
    h, w, na = 6, 9, 9
    sx, sy = np.meshgrid(np.arange(w) * 16, np.arange(h) * 16)
    shifts = np.vstack((sx.ravel(), sy.ravel(), sx.ravel(), sy.ravel())).transpose()
    anchors = (g["frcnn_base_anchors"].reshape((1, na, 4)) + shifts.reshape(1, h * w, 4).transpose((1, 0, 2))).reshape(h * w * na, 4)
    deltas = rs.normal(0, 0.4, size=(h * w * na, 4))
    decoded = fprops.corner_pixels_to_center_inv(anchors, deltas)
    g["frcnn_anchors"], g["frcnn_deltas"], g["frcnn_decoded"] = anchors, deltas, decoded
    scores = rs.permutation(h * w * na).astype(np.float64).reshape(-1, 1) / (h * w * na)
    g["frcnn_scores_all"] = scores.copy()
    with contextlib.redirect_stdout(io.StringIO()):
        fb = fprops.FilterBoxes([96, 144, 3], 16, 10 ** 9, 10 ** 9, 0.7, decoded.copy(), scores.copy())
        fb.clip_boxes()
        g["frcnn_clipped"] = fb.boxes.copy()
        fb.filter_min_size()
        g["frcnn_filter_keep_idx"], g["frcnn_filtered"] = fb.keep_idx.copy(), fb.boxes.copy()
        g["frcnn_filtered_scores"] = fb.scores.copy()
        order = fb.scores.ravel().argsort()[::-1]          # the intended (flattened) ordering
        sorted_boxes = fb.boxes[order]
        g["frcnn_nms_in_sorted"] = sorted_boxes
        for thr in (0.2, 0.7):
            kept = fprops.non_max_suppression(sorted_boxes, fb.scores.ravel()[order].reshape(-1, 1), thr, 50)
            g[f"frcnn_nms_out_thr{int(thr * 10)}"] = kept
    np.savez_compressed(OUT, **g)
    print("wrote", OUT, os.path.getsize(OUT), "bytes,", len(g), "arrays")


if __name__ == "__main__":
    main()
