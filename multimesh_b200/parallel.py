"""
Multi-GPU partitioning of the target points.

The path shards by independent units: every target point's result is a pure function of
(point, source mesh).  The source mesh and its index are replicated on every GPU, the target
points are split into `world` shards, one process per GPU (torchrun).  There is NO collective
inside the compute path; the only communication is the optional final gather of the [N/G, F]
results onto one rank (NCCL on GPUs; gloo in the CPU tests of this host logic).
Results are bit-identical for every G because no arithmetic depends on the partition.

Two partitions:
  "contiguous"  index ranges -- right for target sets that are already spatially ordered (the GLL
                points of a mesh in element order): every range is a compact region.
  "slab"        equal-count slabs along the longest axis -- for unordered point clouds.  A rank then
                touches 1/G of the source elements with the full point density instead of all of them
                at 1/G of the density (each element block is fetched once per warp that needs it, so
                sparse points re-fetch more): 100 M random points on 8 GPUs, 25.6 ms -> see
                profiles/README.md.
"""
from typing import Callable, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(n: int, world: int) -> np.ndarray:
    """world + 1 offsets of contiguous shards whose sizes differ by at most one."""
    base, rem = divmod(int(n), int(world))
    sizes = np.full(world, base, dtype=np.int64)
    sizes[:rem] += 1
    return np.concatenate([[0], np.cumsum(sizes)])


def local_slice(n: int, rank: int, world: int) -> slice:
    b = shard_bounds(n, world)
    return slice(int(b[rank]), int(b[rank + 1]))


def gather_rows(local: torch.Tensor, n_total: int, dst: int = 0, group=None) -> Optional[torch.Tensor]:
    """Gather row-shards (split with `shard_bounds`) onto rank `dst`; other ranks get None.
    Shards are padded to the largest shard so one all_gather_into_tensor / gather suffices."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    b = shard_bounds(n_total, world)
    width = int((b[1:] - b[:-1]).max())
    assert local.shape[0] == int(b[rank + 1] - b[rank]), "shard size does not match shard_bounds"
    pad = torch.zeros((width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    if dist.get_backend(group) == "nccl":
        # NCCL: all_gather into one registered buffer (uniform NVSwitch bandwidth); dst slices it
        out = torch.empty((world * width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, pad, group=group)
        chunks = list(out.split(width)) if rank == dst else None
    else:
        chunks = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
        dist.gather(pad, chunks, dst=dst, group=group)
    if rank != dst:
        return None
    return torch.cat([chunks[r][: int(b[r + 1] - b[r])] for r in range(world)], dim=0)


def slab_partition(points: np.ndarray, world: int, axis: Optional[int] = None):
    """Index sets of `world` equal-count slabs along `axis` (default: the longest extent).  Deterministic:
    every rank computes the same sets from the same array.  Ties on a slab boundary stay together."""
    x = np.asarray(points)
    if axis is None:
        axis = int(np.argmax(x.max(axis=0) - x.min(axis=0))) if len(x) else 0
    if world == 1 or len(x) == 0:
        return [np.arange(len(x))] + [np.arange(0)] * (world - 1)
    cuts = np.quantile(x[:, axis], np.linspace(0.0, 1.0, world + 1)[1:-1])
    bucket = np.searchsorted(cuts, x[:, axis], side="right")
    return [np.flatnonzero(bucket == r) for r in range(world)]


def gather_indexed(local: torch.Tensor, index_sets, dst: int = 0, group=None) -> Optional[torch.Tensor]:
    """Gather shards of arbitrary sizes onto rank `dst` and place shard r at rows index_sets[r]."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n_total = int(sum(len(ix) for ix in index_sets))
    width = max(1, max(len(ix) for ix in index_sets))
    assert local.shape[0] == len(index_sets[rank]), "shard size does not match its index set"
    pad = torch.zeros((width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    if dist.get_backend(group) == "nccl":
        out = torch.empty((world * width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, pad, group=group)
        chunks = list(out.split(width)) if rank == dst else None
    else:
        chunks = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
        dist.gather(pad, chunks, dst=dst, group=group)
    if rank != dst:
        return None
    full = torch.empty((n_total,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    for r in range(world):
        ix = torch.from_numpy(np.asarray(index_sets[r], dtype=np.int64)).to(local.device)
        full[ix] = chunks[r][: len(index_sets[r])]
    return full


def interpolate_sharded(compute: Callable[[np.ndarray], Tuple[torch.Tensor, ...]], points: np.ndarray,
                        gather_to: Optional[int] = 0, group=None, partition: str = "contiguous"):
    """Run `compute(points_shard)` on this rank's shard of `points` (see the module docstring for the two
    partitions) and optionally gather the first returned tensor (the values), in the original point order,
    onto rank `gather_to`.  Returns (local_outputs, gathered_values_or_None)."""
    if dist.is_available() and dist.is_initialized():
        world, rank = dist.get_world_size(group), dist.get_rank(group)
    else:
        world, rank = 1, 0
    if partition == "contiguous":
        sl = local_slice(points.shape[0], rank, world)
        outs = compute(points[sl])
        gathered = None
        if gather_to is not None:
            gathered = outs[0] if world == 1 else gather_rows(outs[0], points.shape[0], gather_to, group)
        return outs, gathered
    if partition != "slab":
        raise ValueError(f"unknown partition {partition!r}")
    sets = slab_partition(points, world)
    outs = compute(np.ascontiguousarray(points[sets[rank]]))
    gathered = None
    if gather_to is not None:
        if world == 1:
            gathered = torch.empty_like(outs[0])
            gathered[torch.from_numpy(sets[0])] = outs[0]
        else:
            gathered = gather_indexed(outs[0], sets, gather_to, group)
    return outs, gathered
