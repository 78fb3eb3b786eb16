"""Box / anchor utilities with the reference's names (MaskRCNN/building_blocks/utils.py:155-369).

Anchor generation runs on the GPU (``od_gen_anchors``: fp64 per anchor, bit-exact with numpy); the three
box-normalisation helpers are the same tiny host-side numpy arithmetic as in the reference (they prepare a
4-element window / post-process <= 100 detection rows and are not part of the device path).
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


def norm_boxes(box, img_shape):
    """utils.py:181-196: pixel -> normalised coordinates (fp64 divide, float32 result)."""
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
