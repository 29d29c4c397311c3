"""e2e leg of bench.py alone, resident source (mm_source_interpolate_host on pinned host buffers): per-setting wall
times of the chunk schedule.  Usage: python tools/e2e_resident.py"""
import ctypes as C
import os
import sys
import time

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402  (workload generators only)
from multimesh_b200 import _lib, ops  # noqa: E402

# under torchrun: one rank per GPU, all ranks copy at the same time (the host side is shared)
world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(local)
bench.bind_to_gpu_numa(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

w = dict(bench.WORKLOADS["S2"], name="S2")
nodes_h, fields_h = bench.make_source(w)
pts_h = bench.make_targets(w, rank)
E, P = nodes_h.shape[0], nodes_h.shape[1]
N, F = pts_h.shape[0], fields_h.shape[1]
lib = _lib.load_lib()
pin = lambda a: torch.from_numpy(a).pin_memory()  # noqa: E731
nodes_p, fields_p, pts_p = pin(nodes_h), pin(fields_h), pin(pts_h)
vals_p = torch.empty((N, F), dtype=torch.float64).pin_memory()
prm = ops.V1().to_c()
nf = C.c_int64(0)
src = C.c_void_p()
_lib.check(lib.mm_source_create_host(C.byref(src), 2, 3, E, C.c_void_p(nodes_p.data_ptr()), F,
                                     C.c_void_p(fields_p.data_ptr()), 1), "create")
sums = set()
for ramp, chunk in (("0", str(1 << 21)), ("1", None), ("0", str(1 << 21)), ("1", None), ("1", str(1 << 21)),
                    ("0", str(1 << 20)), ("1", str(1 << 22))):
    os.environ["MM_HOST_RAMP"] = ramp
    if chunk:
        os.environ["MM_HOST_CHUNK"] = chunk
    else:
        os.environ.pop("MM_HOST_CHUNK", None)
    ts = []
    for i in range(7):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t = time.perf_counter()
        _lib.check(lib.mm_source_interpolate_host(src, N, C.c_void_p(pts_p.data_ptr()), 20, C.byref(prm),
                                                  C.c_void_p(vals_p.data_ptr()), None, None, C.byref(nf)), "run")
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t) * 1e3)
    sums.add(float(vals_p.sum()))
    print(f"rank {rank}/{world} ramp={ramp} chunk={chunk or 'default'}: min {min(ts[2:]):.2f} median {np.median(ts[2:]):.2f} ms", flush=True)
print("checksums identical:", len(sums) == 1, sums)
lib.mm_source_destroy(src)
if world > 1:
    dist.destroy_process_group()
