"""Box / anchor utilities with the reference's names (MaskRCNN/building_blocks/utils.py:155-369).

Anchor generation runs on the GPU (``od_gen_anchors``: fp64 per anchor, bit-exact with numpy). ``norm_boxes`` /
``norm_boxes_tf`` run on the GPU (``od_norm_boxes``) when they are handed a CUDA tensor; for host inputs
``norm_boxes`` / ``denorm_boxes`` stay the reference's own two-line numpy formulas (a 4-element window).
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib


def get_resnet_stage_shapes(conf, image_shape):
    """utils.py:155-178: feature-map shape of every pyramid stage."""
    return np.array([[int(np.ceil(image_shape[0] / stride)), int(np.ceil(image_shape[1] / stride))]
                     for stride in conf.RESNET_STRIDES])


def _norm_boxes_cuda(box, img_shape, tf_float32):
    h, w = int(img_shape[0]), int(img_shape[1])
    if not (box.dtype in (torch.int32, torch.float32, torch.float64)):
        box = box.to(torch.float32 if box.dtype.is_floating_point else torch.int32)
    box = box.contiguous()
    out = torch.empty(box.shape, dtype=torch.float32, device=box.device)
    dl = _lib.DL()
    _lib.check(_lib.lib().od_norm_boxes(dl(box), h, w, 1 if tf_float32 else 0, dl(out), _lib.stream_ptr(box.device)),
               "od_norm_boxes")
    return out


def norm_boxes_tf(boxes, img_shape) -> torch.Tensor:
    """utils.py:198-210: the in-graph float32 normalisation of the GT boxes (training.py:135),
    (boxes - [0,0,1,1]) / (float32([h,w,h,w]) - 1), every operation in float32, on the GPU. Not bit-identical to
    ``norm_boxes`` (float64 divide, one rounding) - a training caller must use this one to reproduce the
    reference's GT inputs of BuildDetectionTargets."""
    return _norm_boxes_cuda(_lib.as_cuda(boxes, torch.float32), img_shape, True)


def norm_boxes(box, img_shape):
    """utils.py:181-196: pixel -> normalised coordinates (fp64 divide, float32 result). A CUDA tensor is
    normalised on the device (no host round trip); anything else on the host with numpy like the reference."""
    if isinstance(box, torch.Tensor) and box.is_cuda:
        return _norm_boxes_cuda(box, img_shape, False)
    if isinstance(box, torch.Tensor):
        box = box.numpy()
    h, w = img_shape
    scale = np.array([h - 1, w - 1, h - 1, w - 1])
    shift = np.array([0, 0, 1, 1])
    return np.divide((np.asarray(box) - shift), scale).astype(np.float32)


def denorm_boxes(boxes, shape):
    """utils.py:212-227: normalised -> integer pixel coordinates."""
    h, w = shape
    scale = np.array([h - 1, w - 1, h - 1, w - 1])
    shift = np.array([0, 0, 1, 1])
    return np.around(np.multiply(boxes, scale) + shift).astype(np.int32)


def anchor_spec(image_shape, scales, ratios, feature_map_shapes, feature_map_strides, anchor_strides):
    return _lib.make_anchor_spec(image_shape, scales, ratios, feature_map_shapes, feature_map_strides, anchor_strides)


def gen_anchors(image_shape, batch_size, scales, ratios, feature_map_shapes, feature_map_strides, anchor_strides,
                device=None) -> torch.Tensor:
    """utils.py:336-353: [batch, A, 4] float32 normalised anchors (level-major, y, x, ratio fastest) on the GPU."""
    spec = anchor_spec(image_shape, scales, ratios, feature_map_shapes, feature_map_strides, anchor_strides)
    L = _lib.lib()
    A = L.od_anchor_count(ctypes.byref(spec))
    if A < 0:
        _lib.check(-8, "od_anchor_count")
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    out = torch.empty((batch_size, A, 4), dtype=torch.float32, device=device)
    dl = _lib.DL()
    _lib.check(L.od_gen_anchors(ctypes.byref(spec), 1, dl(out), _lib.stream_ptr(device)), "od_gen_anchors")
    return out


def gen_anchors_pixel_coord(scales, ratios, feature_map_shapes, feature_map_strides, anchor_strides,
                            device=None) -> torch.Tensor:
    """utils.py:357-369: [A, 4] float64 pixel-coordinate anchors on the GPU."""
    spec = anchor_spec((2, 2), scales, ratios, feature_map_shapes, feature_map_strides, anchor_strides)
    L = _lib.lib()
    A = L.od_anchor_count(ctypes.byref(spec))
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    out = torch.empty((A, 4), dtype=torch.float64, device=device)
    dl = _lib.DL()
    _lib.check(L.od_gen_anchors(ctypes.byref(spec), 0, dl(out), _lib.stream_ptr(device)), "od_gen_anchors")
    return out
