"""Pins the CPU oracle against golden vectors produced by the reference's own numpy code
(tests/golden/make_golden.py) and the notebook constants G1-G8 of SURVEY.md §4."""
import hashlib

import numpy as np

import oracle

RATIOS = [0.5, 1, 2]
STRIDES = [4, 8, 16, 32, 64]


def test_stage_shapes(golden):
    assert np.array_equal(oracle.get_resnet_stage_shapes(STRIDES, [128, 128, 3]), golden["stage_shapes_128"])
    assert np.array_equal(oracle.get_resnet_stage_shapes(STRIDES, [1024, 1024, 3]), golden["stage_shapes_1024"])
    assert np.array_equal(oracle.get_resnet_stage_shapes(STRIDES, [192, 320, 3]), golden["stage_shapes_192x320"])


def test_anchors_coco_1024(golden):  # G1
    a = oracle.gen_anchors([1024, 1024, 3], 1, (32, 64, 128, 256, 512), RATIOS, golden["stage_shapes_1024"], STRIDES, 1)
    assert a.shape == (1, 261888, 4) and a.dtype == np.float32
    assert np.array_equal(a[0, golden["anchors_1024_sample_rows"]], golden["anchors_1024_sample"])
    assert np.array_equal(np.array([a.min(), a.max()], np.float32), golden["anchors_1024_minmax"])
    digest = np.frombuffer(hashlib.sha256(np.ascontiguousarray(a[0]).tobytes()).digest(), np.uint8)
    assert np.array_equal(digest, golden["anchors_1024_sha256"])          # bit-exact, all 261,888 rows


def test_anchors_toy(golden):  # G2, G3
    pix = oracle.gen_anchors_pixel_coord((8, 16, 32, 64, 128), RATIOS, golden["stage_shapes_128"], STRIDES, 1)
    assert np.array_equal(pix, golden["anchors_toy_pixel"])
    assert abs(pix.min() + 90.5096679919) < 1e-9 and abs(pix.max() - 154.509667992) < 1e-9
    assert np.allclose(pix[[3970, 4054, 4074]],
                       [[64, 32, 96, 64], [0, 64, 64, 128], [50.745166, 41.372583, 141.254834, 86.627417]], atol=1e-6)
    norm = oracle.gen_anchors([128, 128, 3], 2, (8, 16, 32, 64, 128), RATIOS, golden["stage_shapes_128"], STRIDES, 1)
    assert np.array_equal(norm, golden["anchors_toy_norm"])


def test_anchors_rect_stride2(golden):
    args = ((16, 32, 64, 128, 256), [0.5, 1, 2, 3], golden["stage_shapes_192x320"], STRIDES, 2)
    assert np.array_equal(oracle.gen_anchors_pixel_coord(*args), golden["anchors_rect_s2_pixel"])
    assert np.array_equal(oracle.gen_anchors([192, 320, 3], 1, *args), golden["anchors_rect_s2_norm"])


def test_norm_denorm_boxes(golden):  # G4, G5
    assert np.array_equal(oracle.norm_boxes(golden["norm_in_gt"], (128, 128)), golden["norm_out_gt"])
    assert np.array_equal(oracle.norm_boxes(golden["norm_in_window"], (1024, 1024)), golden["norm_out_window"])
    assert np.allclose(golden["norm_out_window"], [0.12805474, 0., 0.87194526, 1.], atol=1e-7)
    assert np.array_equal(oracle.norm_boxes(golden["norm_in_rand"], (800, 1024)), golden["norm_out_rand"])
    assert np.array_equal(oracle.denorm_boxes(golden["denorm_in_rand"], (800, 1024)), golden["denorm_out_rand"])


def test_numpy_nms_cases_agree_with_tf_nms(golden):  # G6
    """utils.non_max_supression (utils.py:43-65) and the TF-NMS restatement must select the same boxes on
    well-formed boxes with distinct scores (float32-exact inputs), since both are greedy IoU>thr NMS."""
    assert golden["npnms_keep_0"].tolist() == [3, 2, 1]
    for case in range(6):
        boxes = golden[f"npnms_boxes_{case}"].astype(np.float32)
        scores = golden[f"npnms_scores_{case}"].astype(np.float32)
        thr = float(golden[f"npnms_thr_{case}"])
        keep = oracle.nms(boxes, scores, boxes.shape[0], thr)
        assert keep.tolist() == golden[f"npnms_keep_{case}"].tolist(), case


def test_numpy_iou_agrees(golden):
    box, others = golden["npiou_box"], golden["npiou_boxes"]
    ref = golden["npiou_out"]
    got = np.array([oracle.tf_iou(box, o) for o in others])
    assert np.allclose(got, ref, rtol=1e-5, atol=1e-7)
    got2 = np.array([oracle.target_iou(box, o) for o in others])
    assert np.allclose(got2, ref, rtol=1e-5, atol=1e-7)


def test_frcnn_pieces(golden):
    assert np.array_equal(oracle.FRCNN_BASE_ANCHORS, golden["frcnn_base_anchors"])
    dec = oracle.frcnn_decode(golden["frcnn_anchors"], golden["frcnn_deltas"])
    assert np.allclose(dec, golden["frcnn_decoded"], rtol=1e-14, atol=1e-12)
    for thr in (0.2, 0.7):
        keep = oracle.frcnn_nms_sorted(golden["frcnn_nms_in_sorted"], thr, 50)
        assert np.array_equal(golden["frcnn_nms_in_sorted"][keep], golden[f"frcnn_nms_out_thr{int(thr * 10)}"])


def test_frcnn_full_layer_matches_reference_pieces(golden):
    """Full intended pipeline == reference decode -> clip -> min-size filter -> sort -> reference NMS."""
    h, w, na = 6, 9, 9
    deltas = golden["frcnn_deltas"]
    all_scores = golden["frcnn_scores_all"]
    probs = np.zeros((1, h, w, 2 * na))
    probs[0, :, :, :na] = all_scores.reshape(h, w, na)
    bbox = deltas.reshape(1, h, w, 4 * na)
    for thr in (0.2, 0.7):
        out = oracle.frcnn_proposals(probs, bbox, 96, 144, 10 ** 9, 50, thr)
        ref = golden[f"frcnn_nms_out_thr{int(thr * 10)}"]
        assert out.shape == (ref.shape[0], 5)
        assert np.array_equal(out[:, 1:], ref.astype(np.float32)) and np.all(out[:, 0] == 0)


def test_rpn_targets_match_the_reference():
    """oracle.rpn_targets against the reference's own build_rpn_targets output (tests/golden/make_golden_rpn.py):
    labels, subsampling (replayed np.random.choice permutations) and box deltas; includes SURVEY G7."""
    import os
    import oracle
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_rpn_targets.npz"))
    for case in ("toy", "sub"):
        pos, cls, bbox, counts = oracle.rpn_targets(g[case + "_anchors"], g[case + "_gt"], g[case + "_perm_pos"],
                                                    g[case + "_perm_neg"], int(g[case + "_max_targets"]), [0.1, 0.1, 0.2, 0.2])
        assert np.array_equal(cls, g[case + "_cls"])
        assert np.array_equal(bbox, g[case + "_bbox"])
        assert np.array_equal(pos, g[case + "_pos_anchors"])
        assert counts[0] == int(g[case + "_n_pos0"]) and counts[1] == int(g[case + "_n_neg0"])
    assert (g["toy_cls"] == 1).sum() == 3 and (g["toy_cls"] == -1).sum() == 253      # G7
