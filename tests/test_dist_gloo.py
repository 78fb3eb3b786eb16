"""world_size-2 gloo tests (CPU) of the image-sharded multi-GPU plumbing: shard ranges and the detection all-gather.
The kernels themselves are per image and are parity-tested on one GPU (test_gpu_parity.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from objectdetection_b200.distributed import gather_detections, shard_batch, shard_range


def test_shard_range_partitions_the_batch():
    for batch in (0, 1, 2, 7, 8, 64):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(batch, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, batch, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rs = np.random.RandomState(5)
        full = torch.from_numpy(rs.random_sample((batch, 100, 6)).astype(np.float32))     # same on every rank
        inputs = {"probs": torch.arange(batch * 3, dtype=torch.float32).reshape(batch, 3),
                  "fmaps": [torch.arange(batch * 2, dtype=torch.float32).reshape(batch, 2)]}
        mine = shard_batch(inputs)
        lo, hi = shard_range(batch, rank, world)
        assert torch.equal(mine["probs"], inputs["probs"][lo:hi]) and torch.equal(mine["fmaps"][0], inputs["fmaps"][0][lo:hi])
        det_local = full[lo:hi].clone()             # stands in for this rank's DetectionLayer output
        got = gather_detections(det_local, batch=batch)
        got2 = gather_detections(det_local)         # shard sizes discovered with a collective
        ok = torch.equal(got, full) and torch.equal(got2, full)
        with open(os.path.join(out_dir, f"rank{rank}.txt"), "w") as f:
            f.write("ok" if ok else f"mismatch {tuple(got.shape)}")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("batch", [8, 5])
def test_gather_detections_world2(tmp_path, batch):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, batch, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert (tmp_path / f"rank{r}.txt").read_text() == "ok"


def test_single_process_is_identity():
    x = torch.zeros(2, 100, 6)
    assert gather_detections(x) is x
