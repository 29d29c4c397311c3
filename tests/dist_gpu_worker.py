"""
Worker of tests/test_multi_gpu_nccl.py: launched with `python -m torch.distributed.run --nproc-per-node G`,
one rank per GPU, NCCL.  Checks (property 8; section 8e) with the REAL kernels:

  * results for G GPUs are bit-identical to the single-GPU result, for the contiguous and the slab partition,
    with the values gathered onto rank 0 by grouped send/recv straight into the rows of the full result
    (rank 0's K3 writes its own rows of that buffer in place);
  * the source mesh can be uploaded by rank 0 alone and broadcast over NCCL to the other ranks;
  * the multi_mesh.api driver gll_2_gll, called by every rank, gives the single-GPU file.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

NAMES = ["QKAPPA", "QMU", "RHO", "VP", "VS"]


def main():
    out_dir = sys.argv[1]
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from multimesh_b200 import meshgen, ops, parallel
    from multimesh_b200.io.store import open_store, write_gll_model

    rng = np.random.default_rng(123)
    nodes_h = meshgen.box_mesh((10, 9, 8), 2, warp=0.03)
    fields_h = meshgen.analytic_fields(nodes_h, NAMES)
    pts_h = np.ascontiguousarray(np.concatenate([rng.uniform(-0.05, 1.05, (60001, 3)),
                                                 nodes_h.reshape(-1, 3)[::11]]))
    # source mesh: rank 0 uploads, the others receive over NCCL
    nodes, fields = parallel.broadcast_source([nodes_h, fields_h] if rank == 0 else [None, None], dev, src=0)
    assert torch.equal(nodes.cpu(), torch.from_numpy(nodes_h)) and torch.equal(fields.cpu(), torch.from_numpy(fields_h))
    E, P, _ = nodes_h.shape
    cent, box = ops.element_geometry(nodes)
    pre = ops.element_presolve(nodes)
    index = ops.GridIndex(nodes.view(E * P, 3)).prepare_sites()
    N = len(pts_h)

    def run(p, out=None):
        return ops.interpolate(index, P, nodes, cent, box, fields, p, 20, ops.V1(), presolve=pre, out=out)

    # single-GPU reference (every rank computes it; only used for comparison)
    ref = run(torch.from_numpy(pts_h).to(dev))
    # contiguous partition, K3 writes rank 0's rows of the gather buffer in place
    b = parallel.shard_bounds(N, world)
    sl = parallel.local_slice(N, rank, world)
    full, mine = parallel.gather_buffer(N, (5,), torch.float64, dev, rank, 0, b)
    res = run(torch.from_numpy(pts_h[sl]).to(dev), out=mine)
    if rank == 0:
        assert res[0].data_ptr() == mine.data_ptr()
    got = parallel.gather_rows(res[0], N, dst=0, full=full)
    elem = parallel.gather_rows(res[1], N, dst=0)
    if rank == 0:
        assert torch.equal(got, ref[0]), "contiguous partition: values differ from the single-GPU result"
        assert torch.equal(elem, ref[1])
    else:
        assert got is None
    # slab partition through interpolate_sharded
    (vals, _), gathered = parallel.interpolate_sharded(
        lambda p: run(torch.from_numpy(np.ascontiguousarray(p)).to(dev))[:2], pts_h, gather_to=0, partition="slab")
    if rank == 0:
        assert torch.equal(gathered, ref[0]), "slab partition: values differ from the single-GPU result"
    # all ranks get everything
    every = parallel.allgather_rows(res[0], N)
    assert torch.equal(every, ref[0])
    # the api driver under torchrun: all ranks call it, rank 0 writes the file
    import multi_mesh.api as api

    tgt = meshgen.box_mesh((7, 7, 6), 2, lo=[0.02] * 3, hi=[0.97] * 3)
    a, t_multi, t_single = (os.path.join(out_dir, n) for n in ("from.npz", "to_multi.npz", "to_single.npz"))
    if rank == 0:
        for path in (t_multi, t_single):
            write_gll_model(path, tgt, np.zeros((tgt.shape[0], 5, 27)), NAMES, np.zeros((tgt.shape[0], 2)),
                            ["fluid", "layer"])
        write_gll_model(a, nodes_h, fields_h, NAMES, np.zeros((E, 2)), ["fluid", "layer"])
    dist.barrier()
    api.gll_2_gll(a, t_multi, parameters="ISO", gradient=True)
    if rank == 0:
        # the single-GPU reference: this rank alone, with the driver blind to the process group (no sharding, and
        # above all no collective -- the other rank is not inside the driver)
        orig_world, orig_barrier = parallel._world, parallel.barrier
        parallel._world = lambda group=None: (1, 0)
        parallel.barrier = lambda group=None: None
        try:
            api.gll_2_gll(a, t_single, parameters="ISO", gradient=True)
        finally:
            parallel._world, parallel.barrier = orig_world, orig_barrier
        with open_store(t_multi, "r") as s1, open_store(t_single, "r") as s2:
            assert np.array_equal(s1.read("MODEL/data"), s2.read("MODEL/data")), "api.gll_2_gll differs under torchrun"
        open(os.path.join(out_dir, "ok"), "w").write(f"world {world}: bit-identical\n")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
