"""CPU-side checks of the C-ABI boundary: the library builds/loads, exports every symbol include/odhead.h declares,
validates arguments before touching CUDA, and the host-side mirror exposes the reference's names."""
import ctypes
import inspect
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def libpath():
    from objectdetection_b200 import build
    return build.build()


def test_every_declared_symbol_is_exported(libpath):
    hdr = open(os.path.join(ROOT, "include", "odhead.h")).read()
    declared = sorted(set(re.findall(r"\b(od_[a-z_0-9]+)\s*\(", hdr)))
    assert len(declared) >= 20
    L = ctypes.CDLL(libpath)
    missing = [n for n in declared if not hasattr(L, n)]
    assert not missing, missing
    from objectdetection_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared          # the ctypes table binds exactly the header's entry points


def test_status_strings_and_version(libpath):
    from objectdetection_b200 import _lib
    L = _lib.lib()
    assert L.od_version() >= 100
    assert L.od_strerror(0) == b"ok"
    for code in range(-8, 0):
        assert L.od_strerror(code) != b"unknown status"
    assert L.od_strerror(-99) == b"unknown status"


def test_header_compiles_as_plain_c(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "odhead.h"\nint main(void){ od_proposal_params p; (void)p; return sizeof(DLTensor) == 48 ? 0 : 1; }\n')
    import subprocess
    exe = tmp_path / "t"
    subprocess.run(["/usr/bin/gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", str(src),
                    "-o", str(exe) + ".o"], check=True)


def test_struct_layouts_match_header(tmp_path):
    """sizeof of every parameter struct as seen by C == the ctypes mirror."""
    import subprocess
    from objectdetection_b200 import _lib
    names = ["od_anchor_spec", "od_proposal_params", "od_proposal_debug", "od_target_params", "od_target_debug",
             "od_detection_params", "od_detection_debug", "od_frcnn_params"]
    src = tmp_path / "s.c"
    body = "".join(f'printf("%zu\\n", sizeof({n}));' for n in names)
    src.write_text('#include <stdio.h>\n#include "odhead.h"\nint main(void){' + body + 'return 0;}\n')
    exe = tmp_path / "s"
    subprocess.run(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    sizes = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    mirrors = [_lib.AnchorSpec, _lib.ProposalParams, _lib.ProposalDebug, _lib.TargetParams, _lib.TargetDebug,
               _lib.DetectionParams, _lib.DetectionDebug, _lib.FrcnnParams]
    assert sizes == [ctypes.sizeof(m) for m in mirrors]


def test_cpu_tensors_are_rejected_without_touching_cuda(libpath):
    """No CPU fallback: a host tensor is refused with OD_ERR_DEVICE during argument validation."""
    torch = pytest.importorskip("torch")
    from objectdetection_b200 import _lib
    L = _lib.lib()
    dl = _lib.DL()
    t = torch.zeros((1, 4, 4))
    assert L.od_apply_box_deltas(dl(t), dl(t), dl(t), None) == -4
    assert b"no CPU path" in L.od_last_error_detail()
    assert L.od_clip_boxes(None, None, None, None) == -1
    i = torch.zeros((1, 4), dtype=torch.int32)
    assert L.od_topk(dl(t[0]), 2, None, dl(i), None, 0, None) == -4
    if not torch.cuda.is_available():
        with pytest.raises(_lib.OdHeadError):
            _lib.as_cuda(np.zeros(3, np.float32), torch.float32)


def test_anchor_count_is_host_side(libpath):
    from objectdetection_b200 import _lib, utils
    from objectdetection_b200.config import config
    c = config()
    spec = utils.anchor_spec(c.IMAGE_SHAPE, c.RPN_ANCHOR_SCALES, c.RPN_ANCHOR_RATIOS,
                             utils.get_resnet_stage_shapes(c, c.IMAGE_SHAPE), c.RESNET_STRIDES, c.RPN_ANCHOR_STRIDE)
    assert _lib.lib().od_anchor_count(ctypes.byref(spec)) == 261888


def test_workspace_queries(libpath):
    from objectdetection_b200 import _lib
    L = _lib.lib()
    p = _lib.ProposalParams((ctypes.c_float * 4)(0.1, 0.1, 0.2, 0.2), 6000, 1000, 0.7)
    n1 = L.od_proposal_workspace_bytes(2, 261888, ctypes.byref(p))
    n2 = L.od_proposal_workspace_bytes(4, 261888, ctypes.byref(p))
    assert 0 < n1 < n2 < (1 << 32)
    assert L.od_nms_workspace_bytes(1, 6000) > 6000 * 94 * 8
    assert L.od_topk_workspace_bytes(2, 261888, 6000) > 0
    assert L.od_detection_workspace_bytes(2, 1000, 81) > 0
    assert L.od_detection_target_workspace_bytes(8, 2000, 100) > 0


def test_layer_signatures_match_reference():
    """Constructor argument names/order of the four layers."""
    import objectdetection_b200 as od
    from objectdetection_b200 import fasterrcnn
    sig = lambda c: list(inspect.signature(c.__init__).parameters)[1:]
    assert sig(od.Proposals)[:7] == ["conf", "batch_size", "rpn_class_probs", "rpn_bbox", "inp_anchors", "training", "DEBUG"]
    assert sig(od.MaskRCNN) == ["image_shape", "pool_shape", "num_classes", "levels", "proposals", "feature_maps", "type", "DEBUG"]
    assert sig(od.BuildDetectionTargets)[:5] == ["conf", "proposals", "gt_class_ids", "gt_bboxes", "DEBUG"]
    assert sig(od.DetectionLayer) == ["conf", "image_shape", "num_batches", "window", "proposals", "mrcnn_class_probs", "mrcnn_bbox", "DEBUG"]
    assert sig(fasterrcnn.Proposals)[:3] == ["mode", "rpn_box_class_prob", "rpn_bbox"]
    for cls, methods in ((od.Proposals, ["get_proposals", "get_proposal_graph", "get_anchors_delta_clipped", "debug_outputs"]),
                         (od.MaskRCNN, ["roi_pooling", "get_pooled_rois", "debug_outputs"]),
                         (od.BuildDetectionTargets, ["get_target_rois", "debug_outputs", "build_detection_target"]),
                         (od.DetectionLayer, ["get_detections", "debug_outputs"])):
        for m in methods:
            assert callable(getattr(cls, m))
    from objectdetection_b200 import detection, proposals, utils
    for mod, names in ((proposals, ["apply_box_deltas", "clip_boxes_to_01"]), (detection, ["unmold_detection"]),
                       (utils, ["gen_anchors", "gen_anchors_pixel_coord", "norm_boxes", "denorm_boxes", "get_resnet_stage_shapes"])):
        for n in names:
            assert callable(getattr(mod, n))


def test_host_utils_golden(golden):
    from objectdetection_b200 import utils
    from objectdetection_b200.config import config
    assert np.array_equal(utils.norm_boxes(golden["norm_in_window"], (1024, 1024)), golden["norm_out_window"])
    assert np.array_equal(utils.norm_boxes(golden["norm_in_rand"], (800, 1024)), golden["norm_out_rand"])
    assert np.array_equal(utils.denorm_boxes(golden["denorm_in_rand"], (800, 1024)), golden["denorm_out_rand"])
    assert np.array_equal(utils.get_resnet_stage_shapes(config, [128, 128, 3]), golden["stage_shapes_128"])


def test_oracle_unmold_golden(golden):
    """unmold_detection is a device kernel in the package (tests/test_gpu_parity*.py); its numpy restatement in the
    oracle is pinned here against the reference-run golden."""
    import oracle
    b, c, s = oracle.unmold_detection((600, 800, 3), (1024, 1024, 3), golden["unmold_in"], np.array([131, 0, 893, 1024]))
    assert np.array_equal(b, golden["unmold_boxes"]) and np.array_equal(c, golden["unmold_class_ids"])
    assert np.array_equal(s, golden["unmold_scores"])


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "objectdetection_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", text, re.M), f
                assert "liboracle" not in text, f
