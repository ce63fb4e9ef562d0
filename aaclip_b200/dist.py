"""Data-parallel sharding of an image batch over the GPUs of one box.

Every image is an independent unit (per-sample LayerNorm / attention, shared anchors), so ranks take
contiguous shards of the batch, keep their anomaly maps local and exchange nothing on the data path; the
only collective is one all-gather of the per-image scores (B/R floats per rank) over NCCL / NVLink, which
the reference has no counterpart for (it is single-device: test.py:140-141).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [begin, end) of `total` items owned by `rank`; the first total % world ranks get one extra."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank {rank} / world {world}")
    base, extra = divmod(total, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def gather_scores(local: torch.Tensor, total: int, group: Optional[dist.ProcessGroup] = None,
                  out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """All-gather per-image values of a sharded batch into one [total, ...] tensor on every rank.

    `local` holds this rank's shard_range(total, rank, world) rows: scores [n], or any [n, ...] per-image record (the
    (min, max) map extrema [n, 2]).  Even shards (total % world == 0, what the benchmark and any fixed-batch loop use)
    are ONE collective straight into `out` (a preallocated [total, ...] tensor; allocated here when not given) - no
    padding, no concatenation, no extra kernels on the step.  Uneven shards are padded to the largest shard so that a
    single fixed-size all_gather (NCCL on GPUs, gloo in the CPU tests) suffices.
    """
    tail = tuple(local.shape[1:])
    if not dist.is_available() or not dist.is_initialized():
        if local.shape[0] != total:
            raise ValueError("single-process gather needs the full batch")
        return local
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    b, e = shard_range(total, rank, world)
    if local.shape[0] != e - b:
        raise ValueError(f"rank {rank} holds {local.shape[0]} rows, expected {e - b}")
    if total % world == 0:
        if out is None:
            out = torch.empty((total,) + tail, dtype=local.dtype, device=local.device)
        elif (tuple(out.shape) != (total,) + tail or out.dtype != local.dtype or out.device != local.device
              or not out.is_contiguous()):
            raise ValueError("out must be a contiguous [total, ...] tensor of local's dtype on local's device")
        dist.all_gather_into_tensor(out.view(-1), local.contiguous().view(-1), group=group)
        return out
    width = -(-total // world)
    padded = torch.zeros((width,) + tail, dtype=local.dtype, device=local.device)
    padded[: e - b] = local
    buf = torch.empty((world * width,) + tail, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(buf.view(-1), padded.view(-1), group=group)
    parts = []
    for r in range(world):
        rb, re_ = shard_range(total, r, world)
        parts.append(buf[r * width: r * width + (re_ - rb)])
    return torch.cat(parts)


def gather_extrema(local: torch.Tensor, total: int, group: Optional[dist.ProcessGroup] = None,
                   out: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """The optional second collective of SURVEY 8(e): per-image map extrema (min, max) [n, 2] of this rank's shard ->
    ([total, 2] on every rank, the global (min, max) [2]).  `metrics_eval` normalises the maps of a whole class set by its
    global minimum and maximum and takes each image's normalised maximum (forward_utils.py:241-254): with the extrema of
    every image on every rank that needs no pixel to leave the GPU that produced it."""
    if local.dim() != 2 or local.shape[1] != 2:
        raise ValueError("extrema must be [n, 2] (min, max) per image")
    allx = gather_scores(local, total, group, out)
    return allx, torch.stack([allx[:, 0].min(), allx[:, 1].max()])
