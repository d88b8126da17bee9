"""Multi-GPU layer: images are independent, so ranks take a static round-robin shard of the image list -- the
same partition the reference writes into its per-job submission files
(/root/reference/diffmining/typicality/compute.py:336-341; scripts/parallel.sh:33) -- every rank holds a full
weight replica, and the only exchange is ONE all-gather of the per-image T maps at the end.
Raw [N,n_cond,4,h,w] grids stay rank-local (they are the per-image .npy files)."""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist


def shard_indices(n_items: int, world_size: int, rank: int) -> List[int]:
    """items rank, rank+world, ... (compute.py:339: `lines[i::sub_split]`)"""
    return list(range(rank, n_items, world_size))


def gather_tmaps(local: torch.Tensor, n_total: int, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """All-gather per-image maps.  `local` = [n_local, ...] for the images `shard_indices(n_total, world, rank)`
    in that order; returns [n_total, ...] in ORIGINAL image order on every rank.  One collective
    (all_gather_into_tensor; NCCL over NVLink on GPUs, gloo in CPU tests); shards are padded to equal length."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        assert local.shape[0] == n_total
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    per = (n_total + world - 1) // world
    assert local.shape[0] == len(shard_indices(n_total, world, rank))
    pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((world * per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad.contiguous(), group=group)
    out = out.view((world, per) + tuple(local.shape[1:]))
    # rank r, position j  <->  image j*world + r
    full = out.transpose(0, 1).reshape((per * world,) + tuple(local.shape[1:]))
    return full[:n_total].contiguous()


def run_sharded(n_images: int, compute_local, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """compute_local(indices) -> [len(indices), ...] maps for this rank's shard; returns the gathered [n_images, ...]."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    idx = shard_indices(n_images, world, rank)
    return gather_tmaps(compute_local(idx), n_images, group)
