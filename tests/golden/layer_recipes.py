"""Seeded INPUT recipes of the layer goldens (tests/golden/reference_layers.npz).

Shared by tests/golden/make_golden_layers.py (which runs the reference's own layer classes on these inputs, here in
the build container) and by the tests (which regenerate the same inputs from the seeds — the legacy MT19937 stream of
np.random.seed / RandomState is stable across numpy versions — and compare the oracle and the CUDA path with the
frozen outputs; nothing here needs /root/reference).

The first three recipes are the reference's own `debug()` input recipes (fixed seed + shapes):
  proposals_tf.py:331-345 (seed 325), maskrcnn.py:327-345 (seed 255), detection.py:285-310 (seed 863).
"""
import numpy as np

f32 = np.float32


class RefConfig:
    """MaskRCNN/config.py:5-62 — the attributes the four layers read."""
    IMAGE_SHAPE = [1024, 1024, 3]
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
    MRCNN_TRAIN_ROIS_PER_IMAGE = 200
    MAX_GT_OBJECTS = 100


class ToyConfig(RefConfig):
    """MaskRCNN/shapes.py:17-48 — the toy "shapes" dataset overrides (image 128^2, anchor scales 8..128, 32 training
    ROIs), plus smaller post-NMS caps and a lower score threshold chosen HERE so that the caps bind on small inputs
    (plain attribute overrides, the way shapes.py itself overrides config.py)."""
    IMAGE_SHAPE = [128, 128, 3]
    RPN_ANCHOR_SCALES = (8, 16, 32, 64, 128)
    MRCNN_TRAIN_ROIS_PER_IMAGE = 32
    POST_NMS_ROIS_TRAINING = 200
    POST_NMS_ROIS_INFERENCE = 100
    DETECTION_MIN_THRESHOLD = 0.5


def _softmax(logits):
    e = np.exp(logits - logits.max(-1, keepdims=True))
    return (e / e.sum(-1, keepdims=True)).astype(f32)


# ---------------------------------------------------------------------------------- Proposals (proposals_tf.py)
def proposals_debug325():
    """proposals_tf.py:331-345 verbatim: seed 325, B=1, 4092 anchors, every input uniform [0,1)."""
    np.random.seed(325)
    probs = np.array(np.random.random((1, 4092, 2)), dtype="float32")
    bbox = np.array(np.random.random((1, 4092, 4)), dtype="float32")
    anchors = np.array(np.random.random((1, 4092, 4)), dtype="float32")
    return dict(conf=RefConfig, batch=1, training=False, probs=probs, bbox=bbox, anchors=anchors)


def _rpn_like(rs, B, A):
    fg = rs.beta(0.5, 4, size=(B, A)).astype(f32)
    probs = np.stack([f32(1) - fg, fg], axis=2).astype(f32)
    bbox = rs.normal(0, 1, size=(B, A, 4)).astype(f32)
    return probs, bbox


def proposals_toy(anchors_fn):
    """Toy config (128^2, 4092 real anchors), B=2, training=True (200 post-NMS) — config 1 of BASELINE.json."""
    rs = np.random.RandomState(11)
    anchors = anchors_fn(ToyConfig, 2)
    probs, bbox = _rpn_like(rs, 2, anchors.shape[1])
    probs[0, 100:140, 1] = probs[0, 100, 1]          # score ties: top_k / NMS must order them by index
    probs[0, 100:140, 0] = f32(1) - probs[0, 100, 1]
    return dict(conf=ToyConfig, batch=2, training=True, probs=probs, bbox=bbox, anchors=anchors)


def proposals_coco(anchors_fn, training=False):
    """COCO shape (config 2 / 3): 261,888 real anchors @1024^2, 6000 pre-NMS -> 1000 (inference) / 2000 (training)."""
    rs = np.random.RandomState(2024 + int(training))
    anchors = anchors_fn(RefConfig, 1)
    probs, bbox = _rpn_like(rs, 1, anchors.shape[1])
    return dict(conf=RefConfig, batch=1, training=training, probs=probs, bbox=bbox, anchors=anchors)


# ---------------------------------------------------------------------------------- MaskRCNN.roi_pooling
def roi_pooling_debug255():
    """maskrcnn.py:327-345 verbatim: seed 255, B=2, P2..P5 uniform [0,1) with D=256, proposals uniform [0,1)^4
    (half of them flipped: NaN / INT_MIN level path, extrapolated rows)."""
    np.random.seed(255)
    fmaps = [np.array(np.random.random((2, s, s, 256)), dtype="float32") for s in (256, 128, 64, 32)]
    proposals = np.array(np.random.random((2, 1000, 4)), dtype="float32")
    return dict(image_shape=[1024, 1024, 3], pool_shape=[7, 7], levels=[2, 3, 4, 5], fmaps=fmaps, proposals=proposals)


def roi_pooling_small(pool):
    """Small pyramid (D=8) with well-formed ROIs, zero-padded rows and ROIs on every level; full output stored."""
    rs = np.random.RandomState(77 + pool)
    fmaps = [rs.random_sample((2, s, s, 8)).astype(f32) for s in (64, 32, 16, 8)]
    n = 60
    s = np.exp(rs.uniform(np.log(4), np.log(250), size=(2, n)))
    r = np.exp(rs.uniform(np.log(0.5), np.log(2.0), size=(2, n)))
    h, w = s / np.sqrt(r), s * np.sqrt(r)
    cy, cx = rs.uniform(0, 256, size=(2, n)), rs.uniform(0, 256, size=(2, n))
    boxes = np.stack([cy - h / 2, cx - w / 2, cy + h / 2, cx + w / 2], axis=2)
    boxes = (np.clip(boxes, 0, 255) / 255).astype(f32)
    boxes[:, -6:] = 0                                   # zero-padded proposals
    boxes[0, 3] = [0.2, 0.3, 0.2, 0.6]                  # zero height
    boxes[1, 5] = [0.0, 0.0, 1.0, 1.0]                  # whole image
    boxes[1, 7] = [0.5, 0.5, 1.25, 1.5]                 # partly outside: extrapolated bins
    return dict(image_shape=[1024, 1024, 3], pool_shape=[pool, pool], levels=[2, 3, 4, 5], fmaps=fmaps, proposals=boxes)


# ---------------------------------------------------------------------------------- BuildDetectionTargets
def _rois(rs, n, image, lo, hi):
    s = np.exp(rs.uniform(np.log(lo), np.log(hi), size=n))
    r = np.exp(rs.uniform(np.log(0.5), np.log(2.0), size=n))
    h, w = s / np.sqrt(r), s * np.sqrt(r)
    cy, cx = rs.uniform(0, image, size=n), rs.uniform(0, image, size=n)
    b = np.stack([cy - h / 2, cx - w / 2, cy + h / 2, cx + w / 2], axis=1)
    return (np.clip(b, 0, image - 1) / (image - 1)).astype(f32)


def targets_case(conf, n_prop, n_pad, n_gt_valid, seed, image):
    """One image of config 3: proposals with trailing zero pads, GT = jittered copies of some proposals (so that
    IoU >= 0.5 positives exist), zero-padded to MAX_GT_OBJECTS, explicit shuffles."""
    rs = np.random.RandomState(seed)
    props = _rois(rs, n_prop, image, 16, image / 2)
    props[n_prop - n_pad:] = 0
    G = conf.MAX_GT_OBJECTS
    gt = np.zeros((G, 4), f32)
    cls = np.zeros((G,), np.int32)
    src = rs.choice(n_prop - n_pad, n_gt_valid, replace=False)
    gt[:n_gt_valid] = props[src] + rs.normal(0, 0.01, size=(n_gt_valid, 4)).astype(f32)
    cls[:n_gt_valid] = rs.randint(1, 81, n_gt_valid)
    # a cluster of proposals around the first GT boxes so that there are more positives than the 33 % quota
    k = min(n_gt_valid, 8)
    extra = np.repeat(gt[:k], 12, axis=0) + rs.normal(0, 0.004, size=(12 * k, 4)).astype(f32)
    props[:extra.shape[0]] = extra
    return dict(conf=conf, proposals=props, gt_class_ids=cls, gt_bboxes=gt,
                perm_pos=rs.permutation(n_prop).astype(np.int32), perm_neg=rs.permutation(n_prop).astype(np.int32))


def targets_cfg3(i):
    return targets_case(RefConfig, 2000, 150 + 40 * i, (5, 37, 100)[i], 300 + i, 1024)


def targets_toy():
    return targets_case(ToyConfig, 200, 20, 3, 310, 128)


def targets_few_positives():
    """Fewer positives than the quota: neg_cnt = int32(float32(1/0.33) * float32(pos_count)) - pos_count matters."""
    d = targets_case(RefConfig, 400, 30, 2, 320, 1024)
    d["proposals"][:96] = _rois(np.random.RandomState(321), 96, 1024, 16, 512)    # remove the positive cluster
    d["proposals"][7] = d["gt_bboxes"][0]
    d["proposals"][19] = d["gt_bboxes"][1]
    d["proposals"][33] = d["gt_bboxes"][1] + f32(0.002)
    return d


# ---------------------------------------------------------------------------------- DetectionLayer
def detection_debug863():
    """detection.py:285-310 verbatim: seed 863, B=1, 8 proposals, 4 classes, window [131,0,893,1024]."""
    np.random.seed(863)
    proposals = np.array(np.random.random((1, 8, 4)), dtype="float32")
    probs = np.array(np.random.random((1, 8, 4)), dtype="float32")
    bbox = np.array(np.random.random((1, 8, 4, 4)), dtype="float32")
    return dict(conf=RefConfig, image_shape=[1024, 1024, 3], window=np.array([[131, 0, 893, 1024]], dtype="int32"),
                proposals=proposals, probs=probs, bbox=bbox)


def detection_coco():
    """COCO shape (config 2): B=2, 1000 ROIs, 81 classes; 20 % of the rows confidently foreground."""
    rs = np.random.RandomState(4242)
    B, N, C = 2, 1000, 81
    proposals = np.stack([_rois(rs, N, 1024, 16, 512) for _ in range(B)])
    proposals[:, 950:] = 0
    logits = rs.normal(0, 3, size=(B, N, C))
    rows = rs.random_sample((B, N)) < 0.2
    cls = rs.randint(1, C, size=(B, N))
    bi, ni = np.nonzero(rows)
    logits[bi, ni, cls[bi, ni]] += 12
    probs = _softmax(logits)
    probs[0, 10:14] = probs[0, 10]                     # identical rows: equal scores, same class
    bbox = rs.normal(0, 1, size=(B, N, C, 4)).astype(f32)
    window = np.array([[131, 0, 893, 1024], [0, 96, 1024, 928]], dtype="int32")
    return dict(conf=RefConfig, image_shape=[1024, 1024, 3], window=window, proposals=proposals, probs=probs, bbox=bbox)


def detection_toy_many_per_class():
    """3 foreground classes, 300 ROIs, most of them confident: the per-class cap of 100 and the final top-100 bind."""
    rs = np.random.RandomState(99)
    B, N, C = 1, 300, 4
    proposals = np.stack([_rois(rs, N, 128, 6, 40) for _ in range(B)])
    logits = rs.normal(0, 1, size=(B, N, C))
    cls = rs.randint(1, C, size=(B, N))
    bi, ni = np.nonzero(np.ones((B, N), bool))
    logits[bi, ni, cls[bi, ni]] += 6
    probs = _softmax(logits)
    bbox = rs.normal(0, 0.5, size=(B, N, C, 4)).astype(f32)
    return dict(conf=ToyConfig, image_shape=[128, 128, 3], window=np.array([[0, 0, 128, 128]], dtype="int32"),
                proposals=proposals, probs=probs, bbox=bbox)


# ---------------------------------------------------------------------------------- norm_boxes_tf
def norm_boxes_tf_case():
    rs = np.random.RandomState(5)
    boxes = rs.randint(0, 1025, size=(3, 50, 4)).astype(f32)
    boxes[0, :3] = [[6, 73, 55, 124], [52, 46, 113, 107], [57, 30, 98, 71]]      # SURVEY §4 G4 (toy GT boxes)
    boxes[1] += rs.random_sample((50, 4)).astype(f32)                             # non-integer pixels too
    return dict(boxes=boxes, shapes=[(128, 128), (1024, 1024), (600, 1000)])
