"""Multi-GPU layer: images are independent, so ranks take a static shard of the image list -- round-robin, the same
partition the reference writes into its per-job submission files
(/root/reference/diffmining/typicality/compute.py:336-341; scripts/parallel.sh:33), or balanced by latent area when the
images differ in size -- every rank holds a full weight replica, and the only exchange is ONE all-gather of the
per-image T maps at the end.  Raw [N,n_cond,4,h,w] grids stay rank-local (they are the per-image .npy
files).  NCCL over NVLink on GPUs, gloo in the CPU tests."""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def _world(group=None) -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def shard_indices(n_items: int, world_size: int, rank: int) -> List[int]:
    """items rank, rank+world, ... (compute.py:339: `lines[i::sub_split]`)"""
    return list(range(rank, n_items, world_size))


def shard_by_area(areas: Sequence[int], world_size: int) -> List[List[int]]:
    """Static area-balanced partition for mixed-size image lists (cost of an image ~ its latent area): longest-
    processing-time-first greedy, ties broken by index, so every rank derives the same partition without talking.
    Returns one ascending index list per rank."""
    order = sorted(range(len(areas)), key=lambda i: (-int(areas[i]), i))
    load = [0] * world_size
    out: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += int(areas[i])
    return [sorted(s) for s in out]


class GatherHandle:
    """result of gather_tmaps_async: `.result()` makes the current stream wait for the collective and returns the
    [n_total, ...] maps in original image order"""

    def __init__(self, work, out, keep, world, per, n_total, tail):
        self.work, self.out, self.keep = work, out, keep
        self.world, self.per, self.n_total, self.tail = world, per, n_total, tail
        self._res = None

    def result(self) -> torch.Tensor:
        if self._res is None:
            if self.work is not None:
                self.work.wait()
            out = self.out.view((self.world, self.per) + self.tail)
            # rank r, position j  <->  image j*world + r
            full = out.transpose(0, 1).reshape((self.per * self.world,) + self.tail)
            self._res = full[: self.n_total].contiguous()
            self.keep = None
        return self._res


def gather_tmaps_async(local: torch.Tensor, n_total: int, group: Optional[dist.ProcessGroup] = None) -> GatherHandle:
    """All-gather per-image maps without blocking the compute stream.  `local` = [n_local, ...] for the images
    `shard_indices(n_total, world, rank)` in that order.  One collective (all_gather_into_tensor); shards are padded to
    equal length."""
    world, rank = _world(group)
    tail = tuple(local.shape[1:])
    if world == 1:
        assert local.shape[0] == n_total
        return GatherHandle(None, local, None, 1, n_total, n_total, tail)
    per = (n_total + world - 1) // world
    assert local.shape[0] == len(shard_indices(n_total, world, rank))
    pad = torch.zeros((per,) + tail, dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((world * per,) + tail, dtype=local.dtype, device=local.device)
    work = dist.all_gather_into_tensor(out, pad, group=group, async_op=True)
    return GatherHandle(work, out, pad, world, per, n_total, tail)


def gather_tmaps(local: torch.Tensor, n_total: int, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """Blocking form of gather_tmaps_async: returns [n_total, ...] in ORIGINAL image order on every rank."""
    return gather_tmaps_async(local, n_total, group).result()


def run_sharded(n_images: int, compute_local, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """compute_local(indices) -> [len(indices), ...] maps for this rank's shard; returns the gathered [n_images, ...]."""
    world, rank = _world(group)
    idx = shard_indices(n_images, world, rank)
    return gather_tmaps(compute_local(idx), n_images, group)


def run_sharded_mixed(shapes: Sequence[Tuple[int, int]], compute_local: Callable[[List[int]], List[torch.Tensor]],
                      group: Optional[dist.ProcessGroup] = None, device=None) -> List[torch.Tensor]:
    """Mixed-size image list (BASELINE config 4 / geo, where images are not rescaled: compute.py:165-180).
    `shapes[i]` = (h, w) of image i's T map; ranks take the area-balanced shards of `shard_by_area`;
    `compute_local(indices)` returns this rank's maps, `maps[k]` of shape shapes[indices[k]] (fp32).  All maps travel in
    ONE all-gather of a flat fp32 buffer padded to the largest shard; returns the list of per-image maps in original
    order on every rank.  Per-image results do not depend on the partition (the engine is batch-invariant), so any world
    size returns the same bits."""
    world, rank = _world(group)
    areas = [int(h) * int(w) for h, w in shapes]
    shards = shard_by_area(areas, world)
    mine = shards[rank]
    maps = compute_local(mine)
    assert len(maps) == len(mine)
    per = max(sum(areas[i] for i in s) for s in shards) if shards else 0
    dev = device if device is not None else (maps[0].device if maps else torch.device("cpu"))
    flat = torch.zeros(max(per, 1), dtype=torch.float32, device=dev)
    off = 0
    for i, m in zip(mine, maps):
        assert tuple(m.shape) == tuple(shapes[i]), f"map {i}: {tuple(m.shape)} != {tuple(shapes[i])}"
        flat[off: off + areas[i]] = m.reshape(-1).to(torch.float32)
        off += areas[i]
    if world > 1:
        out = torch.empty(world * flat.numel(), dtype=torch.float32, device=dev)
        dist.all_gather_into_tensor(out, flat, group=group)
        out = out.view(world, flat.numel())
    else:
        out = flat.view(1, -1)
    res: List[Optional[torch.Tensor]] = [None] * len(shapes)
    for r, s in enumerate(shards):
        off = 0
        for i in s:
            res[i] = out[r, off: off + areas[i]].view(shapes[i]).clone()
            off += areas[i]
    return res  # type: ignore[return-value]
