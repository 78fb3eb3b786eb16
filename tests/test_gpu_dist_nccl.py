"""world_size-2 NCCL test on real GPUs (skipped on a one-GPU box): every rank runs Proposals -> DetectionLayer on ITS
images through the CUDA path, the detections are all-gathered over NCCL, and rank 0 checks that the gathered
[B,100,6] tensor equals the CPU oracle's detections for ALL images, in image order, bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, b_local, out_dir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import bench
        from objectdetection_b200 import DetectionLayer, Proposals, utils
        from objectdetection_b200.config import config
        from objectdetection_b200.distributed import gather_detections, shard_range
        conf = config()
        B = world * b_local
        lo, hi = shard_range(B, rank, world)
        shapes = utils.get_resnet_stage_shapes(conf, conf.IMAGE_SHAPE)
        anchors = utils.gen_anchors(conf.IMAGE_SHAPE, b_local, conf.RPN_ANCHOR_SCALES, conf.RPN_ANCHOR_RATIOS, shapes,
                                    conf.RESNET_STRIDES, conf.RPN_ANCHOR_STRIDE, device=dev)
        A = anchors.shape[1]
        full = bench.synth_rpn_head(77, B, A)                                  # the same global batch on every rank
        mine = {k: torch.from_numpy(np.ascontiguousarray(v[lo:hi])).to(dev) for k, v in full.items()}
        window = np.array([bench.WINDOW_PX] * b_local, np.int32)
        props = Proposals(conf, b_local, mine["probs"], mine["bbox"], anchors).get_proposals()
        det = DetectionLayer(conf, conf.IMAGE_SHAPE, b_local, window, props, mine["hprobs"], mine["hbbox"]).get_detections()
        got = gather_detections(det, batch=B)
        got2 = gather_detections(det)                                          # shard sizes discovered with a collective
        torch.cuda.synchronize()
        msg = "ok"
        if rank == 0:
            o_, conf_, _, anc_, win_ = bench.cpu_setup(B, with_fmaps=False)
            want = bench.cpu_detections(o_, conf_, full, anc_, win_)
            g = got.cpu().numpy()
            if g.shape != want.shape or not np.array_equal(g.view(np.uint32), want.view(np.uint32)):
                msg = f"gathered detections differ from the oracle: {g.shape} vs {want.shape}"
            elif not torch.equal(got, got2):
                msg = "the two gather forms disagree"
            elif not (g[:, :, 4] > 0).any():
                msg = "no detections at all - the check is vacuous"
        with open(os.path.join(out_dir, f"rank{rank}.txt"), "w") as f:
            f.write(msg)
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("b_local", [1, 2])
def test_nccl_gathered_detections_equal_the_oracle(tmp_path, b_local):
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), b_local, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert (tmp_path / f"rank{r}.txt").read_text() == "ok"
