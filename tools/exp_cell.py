"""K1 (first pass of the fused pipeline) vs index cell size, on several source geometries.
Usage (GPU box): python tools/exp_cell.py  -> table of knn1 stage time per MM_INDEX_CELL_SCALE."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from multimesh_b200 import _lib, meshgen, ops  # noqa: E402

dev = torch.device("cuda:0")
lib = _lib.load_lib()


def stage_times(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    prof = C.c_void_p()
    _lib.check(lib.mm_profile_create(C.byref(prof), reps), "create")
    lib.mm_profile_begin(prof)
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    lib.mm_profile_end()
    n = C.c_int(0)
    buf = (C.c_float * (reps * 6))()
    _lib.check(lib.mm_profile_read(prof, C.byref(n), buf), "read")
    lib.mm_profile_destroy(prof)
    return np.array(list(buf)).reshape(reps, 6)[: n.value].mean(axis=0)  # sort, knn1, locate1, rerun, gather, unperm


def case(name, nodes_h, pts, gll_form):
    nodes = torch.from_numpy(nodes_h).to(dev)
    E, P, d = nodes.shape
    cent, box = ops.element_geometry(nodes)
    pre = ops.element_presolve(nodes)
    for sc in SCALES:
        os.environ["MM_INDEX_CELL_SCALE"] = str(sc)
        ix = ops.GridIndex(nodes.view(E * P, d) if gll_form else cent)
        t = stage_times(lambda: ops.interpolate(ix, P if gll_form else 1, nodes, cent, box, None, pts, 20, ops.V1(),
                                                presolve=pre))
        info = ix.info()
        print(f"{name:28s} scale {sc:4.2f}  sort {t[0]:6.3f}  knn1 {t[1]:7.3f}  locate1 {t[2]:7.3f}  rerun {t[3]:6.3f}"
              f"   cell {info['cell_size']:.5g} pts/nonempty {info['M'] / max(info['nonempty_cells'], 1):.2f}",
              flush=True)
        ix.close()


SCALES = [float(s) for s in (sys.argv[1].split(",") if len(sys.argv) > 1 else "0.7,0.8,0.9,1.0,1.1,1.2,1.35,1.5".split(","))]
g = torch.Generator(device=dev)
g.manual_seed(1)
N = 6_000_000
pts = torch.rand((N, 3), dtype=torch.float64, device=dev, generator=g)
case("gll o2 60^3 regular", meshgen.box_mesh((60, 60, 60), 2), pts, True)
case("gll o2 60^3 warped", meshgen.box_mesh((60, 60, 60), 2, warp=0.05), pts, True)
case("gll o4 36^3 regular", meshgen.box_mesh((36, 36, 36), 4), pts, True)
case("centroid o2 100^3 regular", meshgen.box_mesh((100, 100, 100), 2), pts, False)
case("centroid o2 100^3 warped", meshgen.box_mesh((100, 100, 100), 2, warp=0.05), pts, False)
# anisotropic: thin elements (x spacing 4x finer than z)
case("centroid o2 160x80x40", meshgen.box_mesh((160, 80, 40), 2), pts, False)
case("gll o2 96x48x24", meshgen.box_mesh((96, 48, 24), 2), pts, True)
