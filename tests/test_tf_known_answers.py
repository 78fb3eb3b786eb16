"""Known-answer vectors of the third-party dependency the reference's layers call: TensorFlow 1.x
(`tf.image.crop_and_resize`, `tf.image.non_max_suppression`, `tf.nn.top_k`; reference call sites maskrcnn.py:167,
proposals.py:115/191, detection.py:133/157). TensorFlow is not installable in the build container, so the vectors are
restated here from its published kernel tests — tensorflow/core/kernels/crop_and_resize_op_test.cc,
non_max_suppression_op_test.cc — and, for top_k, from the documented tie rule ("if two elements are equal, the
lower-index element appears first"). They pin the oracle's restatement of those kernels (CPU tests) and the CUDA
kernels through the C ABI (GPU tests) to the dependency's own expected outputs."""
import numpy as np
import pytest

import oracle

f32 = np.float32
V = -1.0   # extrapolation value of the "Extrapolated" case

# (image [1,H,W,1] values, H, W, boxes, box_ind, crop (h, w), extrapolation, expected flat values)
CROP_CASES = {
    "2x2To1x1": ([1, 2, 3, 4], 2, 2, [[0, 0, 1, 1]], [0], (1, 1), 0.0, [2.5]),
    "2x2To1x1Flipped": ([1, 2, 3, 4], 2, 2, [[1, 1, 0, 0]], [0], (1, 1), 0.0, [2.5]),
    "2x2To3x3": ([1, 2, 3, 4], 2, 2, [[0, 0, 1, 1]], [0], (3, 3), 0.0, [1, 1.5, 2, 2, 2.5, 3, 3, 3.5, 4]),
    "2x2To3x3Flipped": ([1, 2, 3, 4], 2, 2, [[1, 1, 0, 0]], [0], (3, 3), 0.0, [4, 3.5, 3, 3, 2.5, 2, 2, 1.5, 1]),
    "3x3To2x2": (list(range(1, 10)), 3, 3, [[0, 0, 1, 1], [0, 0, 0.5, 0.5]], [0, 0], (2, 2), 0.0, [1, 3, 7, 9, 1, 2, 4, 5]),
    "3x3To2x2Flipped": (list(range(1, 10)), 3, 3, [[1, 1, 0, 0], [0.5, 0.5, 0, 0]], [0, 0], (2, 2), 0.0,
                        [9, 7, 3, 1, 5, 4, 2, 1]),
    "2x2To3x3Extrapolated": ([1, 2, 3, 4], 2, 2, [[-1, -1, 1, 1]], [0], (3, 3), V, [V, V, V, V, 1, 2, V, 3, 4]),
}

THREE_CLUSTERS = [[0, 0, 1, 1], [0, 0.1, 1, 1.1], [0, -0.1, 1, 0.9], [0, 10, 1, 11], [0, 10.1, 1, 11.1], [0, 100, 1, 101]]
THREE_CLUSTERS_FLIPPED = [[1, 1, 0, 0], [0, 0.1, 1, 1.1], [0, .9, 1, -0.1], [0, 10, 1, 11], [1, 10.1, 0, 11.1], [1, 101, 0, 100]]
SCORES = [.9, .75, .6, .95, .5, .3]
# (boxes, scores, max_output_size, iou_threshold, expected selected indices)
NMS_CASES = {
    "SelectFromThreeClusters": (THREE_CLUSTERS, SCORES, 3, 0.5, [3, 0, 5]),
    "SelectFromThreeClustersFlippedCoordinates": (THREE_CLUSTERS_FLIPPED, SCORES, 3, 0.5, [3, 0, 5]),
    "SelectAtMostTwoBoxesFromThreeClusters": (THREE_CLUSTERS, SCORES, 2, 0.5, [3, 0]),
    "SelectAtMostThirtyBoxesFromThreeClusters": (THREE_CLUSTERS, SCORES, 30, 0.5, [3, 0, 5]),
    "SelectWithNegativeScores": (THREE_CLUSTERS, [s - 10 for s in SCORES], 6, 0.5, [3, 0, 5]),
    "SelectSingleBox": ([[0, 0, 1, 1]], [.9], 3, 0.5, [0]),
    "SelectFromTenIdenticalBoxes": ([[0, 0, 1, 1]] * 10, [.9] * 10, 3, 0.5, [0]),
}


def _crop_inputs(case):
    vals, h, w, boxes, bi, crop, ext, want = CROP_CASES[case]
    img = np.asarray(vals, f32).reshape(1, h, w, 1)
    return img, np.asarray(boxes, f32), np.asarray(bi, np.int32), crop, ext, np.asarray(want, f32)


@pytest.mark.parametrize("case", sorted(CROP_CASES))
def test_oracle_crop_and_resize_known_answers(case):
    img, boxes, bi, crop, ext, want = _crop_inputs(case)
    got = oracle.crop_and_resize(img, boxes, bi, crop[0], crop[1], ext)
    assert got.shape == (boxes.shape[0], crop[0], crop[1], 1)
    assert np.array_equal(got.ravel(), want), (case, got.ravel())


@pytest.mark.parametrize("case", sorted(NMS_CASES))
def test_oracle_nms_known_answers(case):
    boxes, scores, max_out, thr, want = NMS_CASES[case]
    keep = oracle.nms(np.asarray(boxes, f32), np.asarray(scores, f32), max_out, thr)
    assert keep.tolist() == want, (case, keep)


def test_oracle_nms_empty_and_topk_tie_rule():
    assert oracle.nms(np.zeros((0, 4), f32), np.zeros((0,), f32), 3, 0.5).tolist() == []
    vals, idx = oracle.topk(np.asarray([[1, 3, 3, 2, 3, 0]], f32), 4)
    assert idx.tolist() == [[1, 2, 4, 3]] and vals.tolist() == [[3, 3, 3, 2]]


# ------------------------------------------------------------------ the CUDA path, through the C ABI
@pytest.mark.gpu
@pytest.mark.parametrize("case", sorted(CROP_CASES))
def test_cuda_crop_and_resize_known_answers(case):
    from objectdetection_b200.maskrcnn import crop_and_resize
    img, boxes, bi, crop, ext, want = _crop_inputs(case)
    img = np.repeat(img, 4, axis=-1)          # the kernels read float4 channel groups: depth 1 -> 4 equal channels
    got = crop_and_resize(img, boxes, bi, crop, extrapolation_value=ext).cpu().numpy()
    assert got.shape == (boxes.shape[0], crop[0], crop[1], 4)
    for ch in range(4):
        assert np.array_equal(got[..., ch].ravel(), want), (case, ch, got[..., ch].ravel())


@pytest.mark.gpu
@pytest.mark.parametrize("case", sorted(NMS_CASES))
def test_cuda_nms_known_answers(case):
    from objectdetection_b200.proposals import non_max_suppression
    boxes, scores, max_out, thr, want = NMS_CASES[case]
    keep = non_max_suppression(np.asarray(boxes, f32), np.asarray(scores, f32), max_out, thr)
    assert keep.cpu().numpy().tolist() == want, (case, keep)


@pytest.mark.gpu
def test_cuda_topk_tie_rule():
    from objectdetection_b200.proposals import top_k
    vals, idx = top_k(np.asarray([[1, 3, 3, 2, 3, 0]], f32), 4)
    assert idx.cpu().numpy().tolist() == [[1, 2, 4, 3]] and vals.cpu().numpy().tolist() == [[3, 3, 3, 2]]
