"""
World-size-2 test of the multi-GPU host logic on CPU (gloo): target points are partitioned into
contiguous shards, each rank computes its shard independently (here with the CPU oracle standing
in for the CUDA kernels -- tests may use it), results are gathered onto rank 0 and must be
BIT-IDENTICAL to the unsharded result (property 8).
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, tmpdir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from multimesh_b200 import meshgen
    from multimesh_b200.parallel import interpolate_sharded
    from oracle import capi as oracle

    nodes = meshgen.box_mesh((4, 4, 4), 2, warp=0.02)
    fields = meshgen.analytic_fields(nodes, ["VP", "VS", "RHO"])
    pts = np.random.default_rng(7).uniform(-0.02, 1.02, (1001, 3))  # odd size: ragged shards
    cent = oracle.centroids(nodes)

    def compute(p):
        cands = oracle.knn_bruteforce(cent, p, 20)
        elem, xi, _, _ = oracle.locate(2, 3, nodes, p, cands, oracle.V1())
        return torch.from_numpy(oracle.interp(2, 3, fields, elem, xi)), torch.from_numpy(elem)

    (vals, elem), gathered = interpolate_sharded(compute, pts, gather_to=0)
    (vals2, elem2), gathered2 = interpolate_sharded(compute, pts, gather_to=0, partition="slab")
    from multimesh_b200.parallel import slab_partition
    sets = slab_partition(pts, world)
    assert sorted(np.concatenate(sets).tolist()) == list(range(len(pts)))      # a partition ...
    assert abs(len(sets[0]) - len(sets[1])) <= 1                                 # ... of equal counts ...
    ax = int(np.argmax(pts.max(axis=0) - pts.min(axis=0)))
    assert pts[sets[0], ax].max() <= pts[sets[1], ax].min()                      # ... into slabs along the longest axis
    assert vals2.shape[0] == len(sets[rank])
    if rank == 0:
        full_vals, _ = compute(pts)
        assert gathered.shape == full_vals.shape
        assert torch.equal(gathered, full_vals), "sharded result is not bit-identical"
        assert torch.equal(gathered2, full_vals), "slab-sharded result is not bit-identical"
        open(os.path.join(tmpdir, "ok"), "w").write("ok")
    else:
        assert gathered is None and gathered2 is None
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_interpolation_bit_identical_world2(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok")


def test_gather_rows_uneven(tmp_path):
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_gather_worker, args=(3, port, str(tmp_path)), nprocs=3, join=True)
    assert os.path.exists(tmp_path / "ok")


def _gather_worker(rank, world, port, tmpdir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from multimesh_b200.parallel import (allgather_rows, broadcast_source, gather_buffer, gather_rows, local_slice,
                                         shard_bounds)

    full = torch.arange(11 * 4, dtype=torch.float64).reshape(11, 4)
    out = gather_rows(full[local_slice(11, rank, world)], 11, dst=1)
    # every rank gets everything; ragged shards, no padding
    assert torch.equal(allgather_rows(full[local_slice(11, rank, world)].clone(), 11), full)
    # the destination computes its shard in place inside the gather buffer
    buf, mine = gather_buffer(11, (4,), torch.float64, "cpu", rank, 1, shard_bounds(11, world))
    if rank == 1:
        mine.copy_(full[local_slice(11, rank, world)])
        assert torch.equal(gather_rows(mine, 11, dst=1, full=buf), full) and buf.data_ptr() == mine.data_ptr() - mine.storage_offset() * 8
    else:
        assert buf is None and gather_rows(full[local_slice(11, rank, world)], 11, dst=1) is None
    # source mesh: one rank holds it, everybody ends up with it
    a = np.arange(24, dtype=np.float64).reshape(2, 4, 3)
    got = broadcast_source([a, 2 * a] if rank == 0 else [None, None], "cpu", src=0)
    assert np.array_equal(got[0].numpy(), a) and np.array_equal(got[1].numpy(), 2 * a)
    if rank == 1:
        assert torch.equal(out, full)
        open(os.path.join(tmpdir, "ok"), "w").write("ok")
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()
