"""CPU ORACLE (second restatement) — test infrastructure only.

A literal numpy transcription of the reference's TF1 graph code, layer by layer, with
each TensorFlow op replaced by a small pure-numpy emulation (``tf_*`` below).  It is
independent of ``odhead_oracle.c`` (no shared code) so the two restatements check each
other in ``tests/test_oracle_*.py``; it is slow (Python loops) and meant for small cases.
**Parity unpinned** for the layer compositions (see the ``odhead_oracle.c`` header; the TF ops
themselves are anchored to TensorFlow's published kernel-test vectors).

All float arithmetic is numpy float32 (IEEE single, no FMA); exp/log are evaluated in
float64 and rounded to float32 — the same convention as the C oracle and the CUDA kernels.
"""
from __future__ import annotations

import numpy as np

f32 = np.float32


def _exp(x):
    return np.exp(np.asarray(x, np.float64)).astype(f32)


def _log(x):
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.log(np.asarray(x, np.float64)).astype(f32)


# ------------------------------------------------------------------ TF op emulations
def tf_top_k(values, k):
    """tf.nn.top_k(sorted=True) on the last axis: descending, ties -> lower index."""
    values = np.asarray(values)
    idx = np.argsort(-(values + f32(0)) if values.dtype.kind == "f" else -values, axis=-1, kind="stable")[..., :k]
    return np.take_along_axis(values, idx, axis=-1), idx.astype(np.int32)


def tf_iou(bi, bj):
    """IOU of TF's non_max_suppression_op.cc (canonicalised corners, area<=0 -> 0)."""
    bi, bj = np.asarray(bi, f32), np.asarray(bj, f32)
    ymin_i, xmin_i = min(bi[0], bi[2]), min(bi[1], bi[3])
    ymax_i, xmax_i = max(bi[0], bi[2]), max(bi[1], bi[3])
    ymin_j, xmin_j = min(bj[0], bj[2]), min(bj[1], bj[3])
    ymax_j, xmax_j = max(bj[0], bj[2]), max(bj[1], bj[3])
    area_i = (ymax_i - ymin_i) * (xmax_i - xmin_i)
    area_j = (ymax_j - ymin_j) * (xmax_j - xmin_j)
    if area_i <= 0 or area_j <= 0:
        return f32(0)
    iymin, ixmin = max(ymin_i, ymin_j), max(xmin_i, xmin_j)
    iymax, ixmax = min(ymax_i, ymax_j), min(xmax_i, xmax_j)
    inter = max(iymax - iymin, f32(0)) * max(ixmax - ixmin, f32(0))
    return inter / (area_i + area_j - inter)


def tf_non_max_suppression(boxes, scores, max_output_size, iou_threshold):
    boxes, scores = np.asarray(boxes, f32).reshape(-1, 4), np.asarray(scores, f32).reshape(-1)
    order = np.argsort(-(scores + f32(0)), kind="stable")
    thr = f32(iou_threshold)
    selected = []
    for i in order:
        if len(selected) >= max_output_size:
            break
        if all(not (tf_iou(boxes[i], boxes[j]) > thr) for j in reversed(selected)):
            selected.append(int(i))
    return np.array(selected, np.int32)


def tf_crop_and_resize(image, boxes, box_ind, crop_size, extrapolation_value=0.0):
    image = np.asarray(image, f32)
    boxes = np.asarray(boxes, f32).reshape(-1, 4)
    B, H, W, D = image.shape
    ch, cw = crop_size
    out = np.zeros((boxes.shape[0], ch, cw, D), f32)
    ext = f32(extrapolation_value)
    for b, (y1, x1, y2, x2) in enumerate(boxes):
        b_in = int(box_ind[b])
        if b_in < 0 or b_in >= B:
            continue
        hs = (y2 - y1) * f32(H - 1) / f32(ch - 1) if ch > 1 else f32(0)
        ws = (x2 - x1) * f32(W - 1) / f32(cw - 1) if cw > 1 else f32(0)
        for y in range(ch):
            in_y = y1 * f32(H - 1) + f32(y) * hs if ch > 1 else f32(0.5 * float(y1 + y2) * (H - 1))
            if not (in_y >= 0) or not (in_y <= f32(H - 1)):
                out[b, y] = ext
                continue
            top, bot = int(np.floor(in_y)), int(np.ceil(in_y))
            y_lerp = in_y - f32(top)
            for x in range(cw):
                in_x = x1 * f32(W - 1) + f32(x) * ws if cw > 1 else f32(0.5 * float(x1 + x2) * (W - 1))
                if not (in_x >= 0) or not (in_x <= f32(W - 1)):
                    out[b, y, x] = ext
                    continue
                left, right = int(np.floor(in_x)), int(np.ceil(in_x))
                x_lerp = in_x - f32(left)
                tl, tr = image[b_in, top, left], image[b_in, top, right]
                bl, br = image[b_in, bot, left], image[b_in, bot, right]
                t = tl + (tr - tl) * x_lerp
                bo = bl + (br - bl) * x_lerp
                out[b, y, x] = t + (bo - t) * y_lerp
    return out


def _to_int32_x86(x):
    """float32 -> int32 like cvttss2si (NaN / out of range -> INT_MIN)."""
    x = np.asarray(x, f32)
    ok = (x > f32(-2147483904.0)) & (x < f32(2147483648.0))
    return np.where(ok, np.where(ok, x, 0).astype(np.int64), -2 ** 31).astype(np.int64)


# ------------------------------------------------------------------ proposals_tf.py
def apply_box_deltas(pre_nms_anchors, bbox_delta):
    """proposals_tf.py:23-65"""
    a, d = np.asarray(pre_nms_anchors, f32), np.asarray(bbox_delta, f32)
    height = a[:, :, 2] - a[:, :, 0]
    width = a[:, :, 3] - a[:, :, 1]
    center_y = a[:, :, 0] + f32(0.5) * height
    center_x = a[:, :, 1] + f32(0.5) * width
    center_y = center_y + d[:, :, 0] * height
    center_x = center_x + d[:, :, 1] * width
    height = height * _exp(d[:, :, 2])
    width = width * _exp(d[:, :, 3])
    y1 = center_y - f32(0.5) * height
    x1 = center_x - f32(0.5) * width
    y2 = y1 + height
    x2 = x1 + width
    return np.stack([y1, x1, y2, x2], axis=2)


def _mn(a, b):  # Eigen/std::min form: NaN in `a` propagates
    return np.where(b < a, b, a)


def _mx(a, b):
    return np.where(a < b, b, a)


def clip_boxes_to_01(anchor_delta, window):
    """proposals_tf.py:67-94"""
    wy1, wx1, wy2, wx2 = (f32(v) for v in window)
    b = np.asarray(anchor_delta, f32)
    y1 = _mx(_mn(b[..., 0], wy2), wy1)
    x1 = _mx(_mn(b[..., 1], wx2), wx1)
    y2 = _mx(_mn(b[..., 2], wy2), wy1)
    x2 = _mx(_mn(b[..., 3], wx2), wx1)
    return np.stack([y1, x1, y2, x2], axis=-1)


def proposals(conf, rpn_class_probs, rpn_bbox, input_anchors, training=False):
    """Proposals.build, proposals_tf.py:136-214. Returns (proposals, debug dict)."""
    rpn_class_probs, rpn_bbox = np.asarray(rpn_class_probs, f32), np.asarray(rpn_bbox, f32)
    anchors = np.asarray(input_anchors, f32)
    n_after = conf.POST_NMS_ROIS_TRAINING if training else conf.POST_NMS_ROIS_INFERENCE
    scores = rpn_class_probs[:, :, 1]
    bbox_delta = rpn_bbox * np.reshape(np.asarray(conf.RPN_BBOX_STDDEV, f32), [1, 1, 4])
    k = min(conf.PRE_NMS_ROIS_COUNT, anchors.shape[1])
    _, ix = tf_top_k(scores, k)
    bi = np.arange(ix.shape[0])[:, None]
    scores, bbox_delta, anchors = scores[bi, ix], bbox_delta[bi, ix], anchors[bi, ix]
    anchor_delta = apply_box_deltas(anchors, bbox_delta)
    clipped = clip_boxes_to_01(anchor_delta, np.array([0, 0, 1, 1], f32))
    out, keeps = [], []
    for b in range(ix.shape[0]):
        nms_idx = tf_non_max_suppression(clipped[b], scores[b], n_after, conf.RPN_NMS_THRESHOLD)
        p = clipped[b][nms_idx]
        out.append(np.pad(p, [(0, max(n_after - p.shape[0], 0)), (0, 0)]))
        keeps.append(nms_idx)
    dbg = dict(bbox_delta=bbox_delta, ix=ix, scores=scores, anchors=anchors, anchor_delta=anchor_delta,
               anchor_delta_clipped=clipped, keep_idx=keeps)
    return np.stack(out, 0), dbg


# ------------------------------------------------------------------ maskrcnn.py
def roi_pooling(image_shape, pool_shape, levels, proposals_, feature_maps):
    """MaskRCNN.roi_pooling, maskrcnn.py:74-187. Returns (pooled [1,B*N,ph,pw,D], roi_level [B,N])."""
    p = np.asarray(proposals_, f32)
    k0, min_k, max_k = 4, min(levels), max(levels)
    h = p[:, :, 2] - p[:, :, 0]
    w = p[:, :, 3] - p[:, :, 1]
    image_area = f32(image_shape[0] * image_shape[1])
    with np.errstate(divide="ignore", invalid="ignore"):
        v = np.sqrt(h * w) / (f32(224.0) / np.sqrt(image_area))
        roi_level = _log(v) / _log(f32(2.0))
    roi_level = k0 + _to_int32_x86(np.rint(roi_level))            # tf.round: half to even
    roi_level = np.where(roi_level < -2 ** 31, roi_level + 2 ** 32, roi_level)  # int32 wrap
    roi_level = np.minimum(max_k, np.maximum(min_k, roi_level)).astype(np.int32)
    pooled, box_to_level = [], []
    for i, level in enumerate(levels):
        ix = np.argwhere(roi_level == level)                    # tf.where: row-major ascending
        level_boxes = p[ix[:, 0], ix[:, 1]]
        pooled.append(tf_crop_and_resize(feature_maps[i], level_boxes, ix[:, 0], pool_shape))
        box_to_level.append(ix)
    pooled = np.concatenate(pooled, axis=0)
    box_to_level = np.concatenate(box_to_level, axis=0)
    sorting_tensor = box_to_level[:, 0] * 100000 + box_to_level[:, 1]
    _, ix = tf_top_k(sorting_tensor, box_to_level.shape[0])
    ix = ix[::-1]
    return pooled[ix][None], roi_level


# ------------------------------------------------------------------ data_processor.py
def get_iou(prop, gt):
    """BuildDetectionTargets.get_iou_tf, data_processor.py:473-510 -> [n, m]."""
    prop, gt = np.asarray(prop, f32).reshape(-1, 4), np.asarray(gt, f32).reshape(-1, 4)
    p = np.repeat(prop, gt.shape[0], axis=0)
    g = np.tile(gt, (prop.shape[0], 1))
    p_area = (p[:, 2] - p[:, 0]) * (p[:, 3] - p[:, 1])
    g_area = (g[:, 2] - g[:, 0]) * (g[:, 3] - g[:, 1])
    i_y1, i_x1 = _mx(p[:, 0], g[:, 0]), _mx(p[:, 1], g[:, 1])
    i_y2, i_x2 = _mn(p[:, 2], g[:, 2]), _mn(p[:, 3], g[:, 3])
    inter = _mx(i_y2 - i_y1, f32(0)) * _mx(i_x2 - i_x1, f32(0))
    with np.errstate(divide="ignore", invalid="ignore"):
        iou = inter / ((p_area + g_area) - inter)
    return iou.reshape(prop.shape[0], gt.shape[0])


def box_refinement(box, gt_box):
    """box_refinement_tf, data_processor.py:443-471."""
    box, gt_box = np.asarray(box, f32), np.asarray(gt_box, f32)
    height = box[:, 2] - box[:, 0]
    width = box[:, 3] - box[:, 1]
    center_y = box[:, 0] + f32(0.5) * height
    center_x = box[:, 1] + f32(0.5) * width
    gt_height = gt_box[:, 2] - gt_box[:, 0]
    gt_width = gt_box[:, 3] - gt_box[:, 1]
    gt_center_y = gt_box[:, 0] + f32(0.5) * gt_height
    gt_center_x = gt_box[:, 1] + f32(0.5) * gt_width
    with np.errstate(divide="ignore", invalid="ignore"):
        dy = (gt_center_y - center_y) / height
        dx = (gt_center_x - center_x) / width
        dh = _log(gt_height / height)
        dw = _log(gt_width / width)
    return np.stack([dy, dx, dh, dw], axis=1)


def _shuffle(indices, perm):
    """Deterministic stand-in for tf.random_shuffle: order = (p for p in perm if p < len)."""
    order = [int(q) for q in perm if 0 <= q < len(indices)]
    return indices[order]


def build_detection_target(conf, proposals_, gt_class_ids, gt_bboxes, perm_pos, perm_neg):
    """BuildDetectionTargets.build_detection_target, data_processor.py:512-652 (one image)."""
    proposals_, gt_bboxes = np.asarray(proposals_, f32), np.asarray(gt_bboxes, f32)
    gt_class_ids = np.asarray(gt_class_ids, np.int32)
    R = conf.MRCNN_TRAIN_ROIS_PER_IMAGE
    non_zeros = np.sum(np.abs(proposals_), axis=1).astype(bool)
    prop = proposals_[non_zeros]
    nz_gt = gt_class_ids.astype(bool)
    gt_boxes, gt_cls = gt_bboxes[nz_gt], gt_class_ids[nz_gt]
    iou = get_iou(prop, gt_boxes)
    roi_iou_max = iou.max(axis=1) if iou.shape[1] else np.full(iou.shape[0], -np.inf, f32)
    pos_all = np.where(roi_iou_max >= 0.5)[0]
    neg_all = np.where(roi_iou_max < 0.5)[0]
    num_pos_inst = int(R * 0.33)
    pos_indices = _shuffle(pos_all, perm_pos)[:num_pos_inst]
    pos_count = pos_indices.shape[0]
    neg_cnt = int(f32(1 / 0.33) * f32(pos_count)) - pos_count
    neg_indices = _shuffle(neg_all, perm_neg)[:neg_cnt]
    pos_rois, neg_rois = proposals_[pos_indices], proposals_[neg_indices]   # un-compacted gather (:600)
    pos_iou = iou[pos_indices]
    assign = pos_iou.argmax(axis=1) if pos_iou.size else np.zeros(0, np.int64)
    roi_gt_class_ids, roi_gt_boxes = gt_cls[assign], gt_boxes[assign]
    deltas = box_refinement(pos_rois, roi_gt_boxes) / np.asarray(conf.BBOX_STD_DEV, f32)
    rois = np.concatenate([pos_rois, neg_rois], axis=0)
    num_pad = max(R - rois.shape[0], 0)
    rois = np.pad(rois, [(0, num_pad), (0, 0)])
    cls = np.pad(roi_gt_class_ids, [(0, num_pad + neg_rois.shape[0])])
    deltas = np.pad(deltas, [(0, num_pad + neg_rois.shape[0]), (0, 0)])
    dbg = dict(iou=iou, roi_iou_max=roi_iou_max, pos_all=pos_all, neg_all=neg_all, pos_indices=pos_indices,
               neg_indices=neg_indices, pos_count=pos_count, neg_cnt=neg_cnt, assign=assign)
    return rois.astype(f32), cls.reshape(1, -1).astype(np.int32), deltas.astype(f32), dbg


# ------------------------------------------------------------------ detection.py
def detection_layer(conf, window_norm, proposals_, mrcnn_class_probs, mrcnn_bbox):
    """DetectionLayer.build, detection.py:80-260. window_norm = norm_boxes(window) [B,4]."""
    proposals_, probs = np.asarray(proposals_, f32), np.asarray(mrcnn_class_probs, f32)
    mrcnn_bbox, window_norm = np.asarray(mrcnn_bbox, f32), np.asarray(window_norm, f32)
    M = conf.DETECTION_POST_NMS_INSTANCES
    B, N = probs.shape[:2]
    class_ids = probs.argmax(axis=2).astype(np.int32)
    bbox_delta = mrcnn_bbox * np.asarray(conf.BBOX_STD_DEV, f32)
    bi, ni = np.arange(B)[:, None], np.arange(N)[None, :]
    class_scores = probs[bi, ni, class_ids]
    bbox_delta = bbox_delta[bi, ni, class_ids]
    refined = apply_box_deltas(proposals_, bbox_delta)
    detections = []
    for i in range(B):
        clipped = clip_boxes_to_01(refined[i][None], window_norm[i])[0]
        class_id_idx = np.where(class_ids[i] > 0)[0]
        score_id_idx = np.where(class_scores[i] > f32(conf.DETECTION_MIN_THRESHOLD))[0]
        keep_idx = np.intersect1d(class_id_idx, score_id_idx)
        pre_cls, pre_scores, pre_props = class_ids[i][keep_idx], class_scores[i][keep_idx], clipped[keep_idx]
        _, first = np.unique(pre_cls, return_index=True)
        unique_cls = pre_cls[np.sort(first)]                     # tf.unique: first-occurrence order
        post = []
        for c in unique_cls:
            cidx = np.where(pre_cls == c)[0]
            nms_idx = tf_non_max_suppression(pre_props[cidx], pre_scores[cidx], M, conf.DETECTION_NMS_THRESHOLD)
            post.append(keep_idx[cidx[nms_idx]])
        post = np.concatenate(post) if post else np.zeros(0, np.int64)
        post = np.intersect1d(keep_idx, post)                    # set_intersection: ascending
        post_scores = class_scores[i][post]
        num_keep = min(M, post_scores.shape[0])
        _, top = tf_top_k(post_scores, num_keep)
        sel = post[top]
        det = np.concatenate([clipped[sel], class_ids[i][sel].astype(f32).reshape(-1, 1),
                              class_scores[i][sel].reshape(-1, 1)], axis=1)
        detections.append(np.pad(det, [(0, M - det.shape[0]), (0, 0)]))
    return np.stack(detections, 0).astype(f32)
