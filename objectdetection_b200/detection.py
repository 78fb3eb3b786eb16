"""DetectionLayer with the reference's interface (MaskRCNN/building_blocks/detection.py:8-279)."""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib
from .proposals import _stddev4
from .utils import denorm_boxes, norm_boxes


def unmold_detection(original_image_shape, image_shape, detections, image_window):
    """detection.py:8-53 — host post-processing of ONE image's [M,6] detections (numpy, like the reference):
    trim at the first class_id == 0, map from the normalised window back to original-image pixels, drop
    zero-area boxes. Returns (boxes [n,4] int32, class_ids [n] int32, scores [n])."""
    if isinstance(detections, torch.Tensor):
        detections = detections.detach().cpu().numpy()
    image_window = norm_boxes(image_window, image_shape[:2])
    zero_ix = np.where(detections[:, 4] == 0)[0]
    N = zero_ix[0] if zero_ix.shape[0] > 0 else detections.shape[0]
    boxes = detections[:N, :4]
    class_ids = detections[:N, 4].astype(np.int32)
    scores = detections[:N, 5]
    wy1, wx1, wy2, wx2 = image_window
    shift = np.array([wy1, wx1, wy1, wx1])
    wh = wy2 - wy1
    ww = wx2 - wx1
    scale = np.array([wh, ww, wh, ww])
    boxes = np.divide(boxes - shift, scale)
    boxes = denorm_boxes(boxes, original_image_shape[:2])
    exclude_ix = np.where((boxes[:, 2] - boxes[:, 0]) * (boxes[:, 3] - boxes[:, 1]) <= 0)[0]
    if exclude_ix.shape[0] > 0:
        boxes = np.delete(boxes, exclude_ix, axis=0)
        class_ids = np.delete(class_ids, exclude_ix, axis=0)
        scores = np.delete(scores, exclude_ix, axis=0)
    return boxes, class_ids, scores


def unmold_detections_batch(original_image_shapes, image_shape, detections, image_windows):
    """``unmold_detection`` for a whole batch on the device (e.g. right after the multi-GPU all-gather): detections
    [B,M,6] CUDA float32, image_windows [B,4] pixel windows, original_image_shapes [B,2|3] (or one shape for all).
    Returns (boxes [B,M,4] int32, class_ids [B,M] int32, scores [B,M] float32, counts [B] int32); the first
    ``counts[b]`` rows of image b equal the reference function's output, the rest is zero."""
    det = _lib.as_cuda(detections, torch.float32)
    dev = det.device
    B, M = det.shape[0], det.shape[1]
    win = norm_boxes(np.asarray(image_windows).reshape(-1, 4), image_shape[:2])          # detection.py:17, host, tiny
    if win.shape[0] == 1 and B > 1:
        win = np.repeat(win, B, 0)
    shp = np.asarray(original_image_shapes, np.int32).reshape(-1, np.asarray(original_image_shapes).shape[-1])[:, :2]
    if shp.shape[0] == 1 and B > 1:
        shp = np.repeat(shp, B, 0)
    win_t = _lib.const_cuda(np.ascontiguousarray(win, np.float32), torch.float32, dev)
    shp_t = _lib.const_cuda(np.ascontiguousarray(shp, np.int32), torch.int32, dev)
    boxes = torch.empty((B, M, 4), dtype=torch.int32, device=dev)
    cls = torch.empty((B, M), dtype=torch.int32, device=dev)
    scores = torch.empty((B, M), dtype=torch.float32, device=dev)
    counts = torch.empty((B,), dtype=torch.int32, device=dev)
    dl = _lib.DL()
    _lib.check(_lib.lib().od_unmold_detections(dl(det), dl(win_t), dl(shp_t), dl(boxes), dl(cls), dl(scores), dl(counts),
                                               _lib.stream_ptr(dev)), "od_unmold_detections")
    return boxes, cls, scores, counts


class DetectionLayer():
    """Per-ROI argmax class -> class-specific delta decode -> clip to window -> bg/score filter -> per-class NMS
    -> top-100 -> zero-padded [B,100,6] rows (y1,x1,y2,x2,class_id,score). Signature of detection.py:57."""

    def __init__(self, conf, image_shape, num_batches, window, proposals, mrcnn_class_probs, mrcnn_bbox, DEBUG=False):
        self.image_shape = image_shape
        self.num_batches = num_batches
        self.bbox_stddev = conf.BBOX_STD_DEV
        self.detection_post_nms_instances = conf.DETECTION_POST_NMS_INSTANCES
        self.detection_min_thresh = conf.DETECTION_MIN_THRESHOLD
        self.detection_nms_threshold = conf.DETECTION_NMS_THRESHOLD
        self.DEBUG = DEBUG
        if isinstance(window, torch.Tensor):
            window = window.detach().cpu().numpy()
        window = norm_boxes(np.asarray(window), image_shape[:2])   # detection.py:66
        self.window = window
        self.detections = self.build(window, proposals, mrcnn_class_probs, mrcnn_bbox)

    def build(self, window, proposals, mrcnn_class_probs, mrcnn_bbox):
        L = _lib.lib()
        props = _lib.as_cuda(proposals, torch.float32)
        dev = props.device
        probs = _lib.as_cuda(mrcnn_class_probs, torch.float32, dev)
        bbox = _lib.as_cuda(mrcnn_bbox, torch.float32, dev)
        win = _lib.const_cuda(np.asarray(window, np.float32).reshape(-1, 4), torch.float32, dev)
        B, N, C = probs.shape
        if B != self.num_batches:
            raise ValueError(f"num_batches={self.num_batches} but mrcnn_class_probs has batch {B}")
        M = int(self.detection_post_nms_instances)
        params = _lib.DetectionParams(_stddev4(self.bbox_stddev), float(self.detection_min_thresh),
                                      float(self.detection_nms_threshold), M)
        det = torch.empty((B, M, 6), dtype=torch.float32, device=dev)
        dl = _lib.DL()
        dbg = _lib.DetectionDebug()
        if self.DEBUG:
            self.class_ids = torch.empty((B, N), dtype=torch.int32, device=dev)
            self.class_scores = torch.empty((B, N), dtype=torch.float32, device=dev)
            self.bbox_delta = torch.empty((B, N, 4), dtype=torch.float32, device=dev)
            self.refined_proposals = torch.empty((B, N, 4), dtype=torch.float32, device=dev)
            self.clipped_proposals = torch.empty((B, N, 4), dtype=torch.float32, device=dev)
            self.keep_mask = torch.empty((B, N), dtype=torch.int32, device=dev)
            self.nms_keep_mask = torch.empty((B, N), dtype=torch.int32, device=dev)
            dbg.class_ids, dbg.class_scores, dbg.bbox_delta = dl(self.class_ids), dl(self.class_scores), dl(self.bbox_delta)
            dbg.refined_proposals, dbg.clipped_proposals = dl(self.refined_proposals), dl(self.clipped_proposals)
            dbg.keep_mask, dbg.nms_keep_mask = dl(self.keep_mask), dl(self.nms_keep_mask)
        ws = _lib.workspace(L.od_detection_workspace_bytes(B, N, C), dev)
        _lib.check(L.od_detection_forward(dl(props), dl(probs), dl(bbox), dl(win), ctypes.byref(params), dl(det),
                                          ctypes.byref(dbg), ws.data_ptr(), ws.numel(), _lib.stream_ptr(dev)),
                   "od_detection_forward")
        return det

    def get_detections(self):
        return self.detections

    def debug_outputs(self):
        """Same order as detection.py:268-279. The index-plumbing tensors of the TF graph (indices, mesh, ixs) have
        no counterpart; the per-image pre-NMS lists are derived from keep_mask on request."""
        keep = self.keep_mask.bool()
        B = keep.shape[0]
        clipped_list = [self.clipped_proposals[b][None] for b in range(B)]
        pre_cls = [self.class_ids[b][keep[b]] for b in range(B)]
        pre_scores = [self.class_scores[b][keep[b]] for b in range(B)]
        pre_props = [self.clipped_proposals[b][keep[b]] for b in range(B)]
        return (self.class_ids, None, None, None, self.class_scores, self.bbox_delta, self.refined_proposals,
                clipped_list, pre_cls, pre_scores, pre_props)
