"""
Multi-GPU partitioning of the target points.

The path shards by independent units: every target point's result is a pure function of
(point, source mesh).  The source mesh and its index are replicated on every GPU, the target
points are split into `world` shards, one process per GPU (torchrun).  There is NO collective
inside the compute path.  Communication happens only
  * at set-up, optionally: `broadcast_source` ships the source mesh from the rank that loaded it to the
    others over NVLink (NCCL broadcast) instead of every rank pulling its own copy over PCIe;
  * at the end, optionally: `gather_rows` / `gather_indexed` collect the [N/G, F] results on ONE rank
    with grouped point-to-point transfers (ncclSend/ncclRecv via `batch_isend_irecv`) straight into
    the rows of the full result -- the destination's own shard is computed in place (`out_buffer`), so
    there is no padding, no all-gather to ranks that do not want the data and no concatenation copy.
Results are bit-identical for every G because no arithmetic depends on the partition (gloo tests on CPU,
torchrun test on GPUs).

Two partitions:
  "contiguous"  index ranges -- right for target sets that are already spatially ordered (the GLL
                points of a mesh in element order): every range is a compact region.
  "slab"        equal-count slabs along the longest axis -- for unordered point clouds.  A rank then
                touches 1/G of the source elements with the full point density instead of all of them
                at 1/G of the density (each element block is fetched once per warp that needs it, so
                sparse points re-fetch more): 100 M random points on 8 GPUs, 14.7 vs 25.6 ms
                (profiles/README.md).
"""
from typing import Callable, List, Optional, Sequence, Tuple

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


def _world(group=None) -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def barrier(group=None) -> None:
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.barrier(group)


def _global_rank(r: int, group=None) -> int:
    return r if group is None else dist.get_global_rank(group, r)


def gather_buffer(n_total: int, tail_shape: Sequence[int], dtype, device, rank: int, dst: int,
                  bounds: np.ndarray) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
    """(full, mine): on `dst` the full [n_total, ...] result and the view of this rank's rows inside it -- hand
    `mine` to the gather kernel (`ops.interpolate(..., out=mine)`) so that no copy of the local shard is ever
    made; (None, None) on the other ranks, which allocate their own shard."""
    if rank != dst:
        return None, None
    full = torch.empty((int(n_total),) + tuple(tail_shape), dtype=dtype, device=device)
    return full, full[int(bounds[rank]): int(bounds[rank + 1])]


def gather_rows(local: torch.Tensor, n_total: int, dst: int = 0, group=None,
                full: Optional[torch.Tensor] = None) -> Optional[torch.Tensor]:
    """Collect contiguous row-shards (split with `shard_bounds`) on rank `dst`; other ranks get None.
    Every source rank sends its shard once, `dst` receives each shard directly into its rows of the result.
    `full`: the buffer from `gather_buffer` when the local shard already lives inside it."""
    world, rank = _world(group)
    b = shard_bounds(n_total, world)
    assert local.shape[0] == int(b[rank + 1] - b[rank]), "shard size does not match shard_bounds"
    if world == 1:
        return local if full is None else full
    ops = []
    if rank == dst:
        if full is None:
            full = torch.empty((int(n_total),) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        mine = full[int(b[rank]): int(b[rank + 1])]
        if mine.data_ptr() != local.data_ptr():
            mine.copy_(local)
        for r in range(world):
            if r != dst and b[r + 1] > b[r]:
                ops.append(dist.P2POp(dist.irecv, full[int(b[r]): int(b[r + 1])], _global_rank(r, group), group))
    elif local.shape[0] > 0:
        ops.append(dist.P2POp(dist.isend, local.contiguous(), _global_rank(dst, group), group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return full if rank == dst else None


def allgather_rows(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """Every rank ends up with the full [n_total, ...] result (for in-memory target meshes each rank owns a copy
    of): every rank's shard goes straight into its rows of every other rank's result with grouped point-to-point
    transfers; ragged shards need no padding."""
    world, rank = _world(group)
    if world == 1:
        return local
    b = shard_bounds(n_total, world)
    assert local.shape[0] == int(b[rank + 1] - b[rank]), "shard size does not match shard_bounds"
    full = torch.empty((int(n_total),) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    full[int(b[rank]): int(b[rank + 1])].copy_(local)
    ops = []
    src = local.contiguous()
    for r in range(world):
        if r == rank:
            continue
        if src.shape[0] > 0:
            ops.append(dist.P2POp(dist.isend, src, _global_rank(r, group), group))
        if b[r + 1] > b[r]:
            ops.append(dist.P2POp(dist.irecv, full[int(b[r]): int(b[r + 1])], _global_rank(r, group), group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return full


def slab_partition(points: np.ndarray, world: int, axis: Optional[int] = None) -> List[np.ndarray]:
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
    """Collect shards of arbitrary sizes on rank `dst` and place shard r at rows index_sets[r]: received in
    shard order into one staging buffer, then ONE device scatter through the concatenated index."""
    world, rank = _world(group)
    sizes = [len(ix) for ix in index_sets]
    n_total = int(sum(sizes))
    assert local.shape[0] == sizes[rank], "shard size does not match its index set"
    b = np.concatenate([[0], np.cumsum(sizes)])
    staged = None
    ops = []
    if rank == dst:
        staged = torch.empty((n_total,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        staged[int(b[rank]): int(b[rank + 1])].copy_(local)
        for r in range(world):
            if r != dst and sizes[r] > 0:
                ops.append(dist.P2POp(dist.irecv, staged[int(b[r]): int(b[r + 1])], _global_rank(r, group), group))
    elif local.shape[0] > 0 and world > 1:
        ops.append(dist.P2POp(dist.isend, local.contiguous(), _global_rank(dst, group), group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    if rank != dst:
        return None
    order = torch.from_numpy(np.concatenate([np.asarray(ix, dtype=np.int64) for ix in index_sets])).to(local.device)
    full = torch.empty_like(staged)
    full.index_copy_(0, order, staged)
    return full


def broadcast_source(arrays: Sequence[Optional[np.ndarray]], device, src: int = 0, group=None) -> List[torch.Tensor]:
    """Source-mesh arrays (nodes, fields, ...) resident on `device` of EVERY rank while only rank `src` holds them
    on the host: `src` uploads once over PCIe, the other ranks receive over NVLink (NCCL broadcast) -- instead of
    G uploads of the same gigabytes through the host's PCIe/memory system.  `arrays` on the other ranks may be
    None; shapes travel first in a small object broadcast."""
    world, rank = _world(group)
    if world == 1:
        return [torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(device) for a in arrays]
    meta = [None]
    if rank == src:
        meta = [[(tuple(a.shape), str(a.dtype)) for a in arrays]]
    dist.broadcast_object_list(meta, src=_global_rank(src, group), group=group)
    out = []
    for i, (shape, dtype) in enumerate(meta[0]):
        if rank == src:
            t = torch.from_numpy(np.ascontiguousarray(arrays[i])).to(device, non_blocking=True)
        else:
            t = torch.empty(shape, dtype=getattr(torch, dtype), device=device)
        dist.broadcast(t, src=_global_rank(src, group), group=group)
        out.append(t)
    return out


def interpolate_sharded(compute: Callable[[np.ndarray], Tuple[torch.Tensor, ...]], points: np.ndarray,
                        gather_to: Optional[int] = 0, group=None, partition: str = "contiguous"):
    """Run `compute(points_shard)` on this rank's shard of `points` (see the module docstring for the two
    partitions) and optionally gather the first returned tensor (the values), in the original point order,
    onto rank `gather_to`.  Returns (local_outputs, gathered_values_or_None)."""
    world, rank = _world(group)
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
