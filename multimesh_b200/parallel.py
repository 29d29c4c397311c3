"""
Multi-GPU partitioning of the target points.

The path shards by independent units: every target point's result is a pure function of
(point, source mesh).  The source mesh and its index are replicated on every GPU, the target
points are split into `world` contiguous ranges, one process per GPU (torchrun).  There is NO
collective inside the compute path; the only communication is the optional final gather of the
[N/G, F] results onto one rank (NCCL on GPUs; gloo in the CPU tests of this host logic).
Results are bit-identical for every G because no arithmetic depends on the partition.
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


def interpolate_sharded(compute: Callable[[np.ndarray], Tuple[torch.Tensor, ...]], points: np.ndarray,
                        gather_to: Optional[int] = 0, group=None):
    """Run `compute(points_shard)` on this rank's contiguous shard of `points` and optionally gather
    the first returned tensor (the values) onto rank `gather_to`.
    Returns (local_outputs, gathered_values_or_None)."""
    if dist.is_available() and dist.is_initialized():
        world, rank = dist.get_world_size(group), dist.get_rank(group)
    else:
        world, rank = 1, 0
    sl = local_slice(points.shape[0], rank, world)
    outs = compute(points[sl])
    gathered = None
    if gather_to is not None:
        gathered = outs[0] if world == 1 else gather_rows(outs[0], points.shape[0], gather_to, group)
    return outs, gathered
