"""objectdetection_b200 — the detection-head hot path of Sardhendu/ObjectDetection (Mask R-CNN / Faster R-CNN
study code) rebuilt for NVIDIA B200 (sm_100a): hand-written CUDA kernels behind a C ABI (``include/odhead.h``,
``libodhead.so``), driven through layer classes that keep the reference's names, argument order and tensor
layouts. PyTorch is only the tensor carrier (DLPack, streams, torch.distributed). There is no CPU path.

    Proposals                 proposals_tf.py:98     (ProposalLayer)
    MaskRCNN / pyramid_roi_align   maskrcnn.py:74-187     (PyramidROIAlign)
    BuildDetectionTargets     data_processor.py:430  (DetectionTargetLayer)
    DetectionLayer            detection.py:56        (DetectionLayer)
"""
from .config import ShapesConfig, config  # noqa: F401

__all__ = ["config", "ShapesConfig", "Proposals", "MaskRCNN", "BuildDetectionTargets", "DetectionLayer", "Loss"]


def __getattr__(name):
    # torch / CUDA are only imported when a layer is actually requested
    if name == "Proposals":
        from .proposals import Proposals
        return Proposals
    if name == "MaskRCNN":
        from .maskrcnn import MaskRCNN
        return MaskRCNN
    if name == "BuildDetectionTargets":
        from .data_processor import BuildDetectionTargets
        return BuildDetectionTargets
    if name == "DetectionLayer":
        from .detection import DetectionLayer
        return DetectionLayer
    if name == "Loss":
        from .loss_optimize import Loss
        return Loss
    raise AttributeError(name)
