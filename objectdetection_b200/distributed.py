"""Image-sharded multi-GPU plumbing for the detection-head path.

Every layer of the path is per image (the reference loops over the batch in Python: proposals_tf.py:188-196,
detection.py:143, training.py:70), so N GPUs run N independent shards: one process per GPU, a contiguous slice of the
batch per rank, no data-path collective. The only exchange is one ``all_gather`` of the ``[B_local, 100, 6]``
detections (2.4 kB per image) at the end of inference; training targets stay local. ``torch.distributed`` is the
carrier (NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(batch: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous [lo, hi) slice of `batch` images owned by `rank` (earlier ranks take the remainder)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(x, rank: int | None = None, world: int | None = None):
    """This rank's slice along dim 0 of a tensor (or of every tensor of a list / dict)."""
    if rank is None:
        rank, world = dist.get_rank(), dist.get_world_size()
    if isinstance(x, dict):
        return {k: shard_batch(v, rank, world) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return type(x)(shard_batch(v, rank, world) for v in x)
    lo, hi = shard_range(x.shape[0], rank, world)
    return x[lo:hi]


def gather_detections(det_local: torch.Tensor, batch: int | None = None, group=None) -> torch.Tensor:
    """all_gather of per-rank detections ``[B_local, M, 6]`` into ``[B, M, 6]`` in image order on every rank.

    Equal shards use one ``all_gather_into_tensor``; ragged shards (batch not divisible by the world size) are
    padded to the largest shard and trimmed after the gather."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return det_local
    world = dist.get_world_size(group)
    b_local = det_local.shape[0]
    if batch is None:
        sizes = torch.tensor([b_local], dtype=torch.int64, device=det_local.device)
        all_sizes = torch.empty(world, dtype=torch.int64, device=det_local.device)
        dist.all_gather_into_tensor(all_sizes, sizes, group=group)
        counts = [int(v) for v in all_sizes.tolist()]
    else:
        counts = [shard_range(batch, r, world)[1] - shard_range(batch, r, world)[0] for r in range(world)]
    bmax = max(counts)
    src = det_local.contiguous()
    if b_local < bmax:
        pad = torch.zeros((bmax - b_local,) + tuple(det_local.shape[1:]), dtype=det_local.dtype, device=det_local.device)
        src = torch.cat([src, pad], 0)
    out = torch.empty((world * bmax,) + tuple(det_local.shape[1:]), dtype=det_local.dtype, device=det_local.device)
    dist.all_gather_into_tensor(out, src, group=group)
    if all(c == bmax for c in counts):
        return out
    return torch.cat([out[r * bmax:r * bmax + c] for r, c in enumerate(counts)], 0)
