"""Hyper-parameters of the detection-head path, with the reference's attribute names.

Mirrors ``MaskRCNN/config.py:5-62`` and ``MaskRCNN/shapes.py:17-48`` (only the attributes the
hot path reads).  Any object exposing the same attribute names can be passed as ``conf`` to the
layer classes (duck-typed, like the reference).
"""
import numpy as np


class config(object):
    NAME = "coco_heads"
    IMAGE_SHAPE = [1024, 1024, 3]
    NUM_CLASSES = 81

    RESNET_STRIDES = [4, 8, 16, 32, 64]

    RPN_ANCHOR_STRIDE = 1
    RPN_ANCHOR_RATIOS = [0.5, 1, 2]
    RPN_ANCHOR_SCALES = (32, 64, 128, 256, 512)
    RPN_NMS_THRESHOLD = 0.7
    RPN_BBOX_STDDEV = np.array([0.1, 0.1, 0.2, 0.2])
    BBOX_STD_DEV = np.array([0.1, 0.1, 0.2, 0.2])

    PRE_NMS_ROIS_COUNT = 6000
    POST_NMS_ROIS_TRAINING = 2000
    POST_NMS_ROIS_INFERENCE = 1000

    DETECTION_MIN_THRESHOLD = 0.7
    DETECTION_NMS_THRESHOLD = 0.3
    DETECTION_POST_NMS_INSTANCES = 100

    RPN_TRAIN_ANCHORS_PER_IMAGE = 256
    MRCNN_TRAIN_ROIS_PER_IMAGE = 200
    USE_MINI_MASK = True
    MINI_MASK_SHAPE = (56, 56)
    MASK_SHAPE = (28, 28)
    MAX_GT_OBJECTS = 100

    def display(self):
        print("\nConfigurations:")
        for a in dir(self):
            if not a.startswith("__") and not callable(getattr(self, a)):
                print("{:40} {}".format(a, getattr(self, a)))
        print("\n")


class ShapesConfig(config):
    """Toy 128x128 'shapes' dataset override (MaskRCNN/shapes.py:17-48)."""
    NAME = "shapes"
    NUM_CLASSES = 1 + 3
    IMAGE_SHAPE = [128, 128, 3]
    RPN_ANCHOR_SCALES = (8, 16, 32, 64, 128)
    MRCNN_TRAIN_ROIS_PER_IMAGE = 32
