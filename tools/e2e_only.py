"""End-to-end leg of bench.py alone (mm_interpolate_host on pinned host buffers), per-call wall times.
Usage: python tools/e2e_only.py [repo_root]   -- repo_root selects which build of the library is loaded."""
import ctypes as C
import os
import sys
import time

root = os.path.abspath(sys.argv[1]) if len(sys.argv) > 1 else os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root)
import torch  # noqa: E402

here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.append(here)
import bench  # noqa: E402  (workload generators only)
from multimesh_b200 import _lib, ops  # noqa: E402

print("library from", os.path.dirname(_lib.__file__))
w = dict(bench.WORKLOADS["S2"], name="S2")
nodes_h, fields_h = bench.make_source(w)
pts_h = bench.make_targets(w, 0)
E, P = nodes_h.shape[0], nodes_h.shape[1]
N, F = pts_h.shape[0], fields_h.shape[1]
lib = _lib.load_lib()
pin = lambda a: torch.from_numpy(a).pin_memory()  # noqa: E731
nodes_p, fields_p, pts_p = pin(nodes_h), pin(fields_h), pin(pts_h)
vals_p = torch.empty((N, F), dtype=torch.float64).pin_memory()
prm = ops.V1().to_c()
nf = C.c_int64(0)
for i in range(6):
    torch.cuda.synchronize()
    t = time.perf_counter()
    rc = lib.mm_interpolate_host(2, 3, E, C.c_void_p(nodes_p.data_ptr()), F, C.c_void_p(fields_p.data_ptr()), N,
                                 C.c_void_p(pts_p.data_ptr()), 20, 1, C.byref(prm), C.c_void_p(vals_p.data_ptr()),
                                 None, None, C.byref(nf))
    torch.cuda.synchronize()
    print(f"call {i}: {(time.perf_counter() - t) * 1e3:8.2f} ms  rc={rc}", flush=True)
print("checksum", float(vals_p.sum()))
