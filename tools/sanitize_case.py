"""Small end-to-end case for compute-sanitizer (memcheck): all kernel families, odd sizes, array tails."""
import sys
import numpy as np, torch
sys.path.insert(0, '.')
from multimesh_b200 import meshgen, ops
dev = torch.device('cuda:0')
rng = np.random.default_rng(1)
for order, n in ((2, 3), (4, 3), (1, 5)):
    nodes = meshgen.box_mesh((n, n, n), order, warp=0.02)
    E, P, _ = nodes.shape
    fields = rng.normal(size=(E, 5, P))
    pts = np.concatenate([rng.uniform(-0.1, 1.1, (777, 3)), nodes.reshape(-1, 3)[::5]])
    tn, tf, tp = (torch.from_numpy(a).to(dev) for a in (nodes, fields, pts))
    cent, box = ops.element_geometry(tn)
    pre = ops.element_presolve(tn)
    for form in ("gll", "gll_sites", "centroid"):
        index, div = (ops.GridIndex(tn.view(E * P, 3)), P) if form != "centroid" else (ops.GridIndex(cent), 1)
        if form == "gll_sites":
            index.prepare_sites()  # site table: knn_tile_kernel<SITES>, K4's tables
        for spec in (ops.V1(), ops.V2(1.05, True)):
            out, elem, xi, st, nf = ops.interpolate(index, div, tn, cent, box, tf, tp, 20, spec, presolve=pre)
        cands = index.query_idx(tp, 20, divisor=div)
        e, x, s, _ = ops.locate(tn, cent, box, tp, cands, ops.V1(), presolve=pre)
        o = ops.interp(tf, e, x)
        c = ops.coeffs(e, x, order)
    torch.cuda.synchronize()
# K1's sub-tile / give-up paths (tiny staging capacity), K3 with several field passes and many groups per warp, 2-D
import os
for cap, order, dim, F in (("60", 2, 3, 13), ("16", 4, 3, 8), (None, 2, 2, 1), (None, 4, 2, 7), (None, 1, 3, 2)):
    if cap:
        os.environ["MM_KT_CAP"] = cap
    else:
        os.environ.pop("MM_KT_CAP", None)
    nodes = meshgen.box_mesh((4,) * dim, order, warp=0.02)
    E, P, _ = nodes.shape
    fields = rng.normal(size=(E, F, P))
    pts = np.concatenate([rng.uniform(0.0, 0.5, (900, dim)), rng.uniform(-0.1, 1.1, (300, dim))])
    tn, tf, tp = (torch.from_numpy(a).to(dev) for a in (nodes, fields, pts))
    cent, box = ops.element_geometry(tn)
    pre = ops.element_presolve(tn)
    ops.interpolate(ops.GridIndex(cent), 1, tn, cent, box, tf, tp, 20, ops.V1(), presolve=pre)
    ops.interpolate(ops.GridIndex(tn.view(E * P, dim)).prepare_sites(), P, tn, cent, box, tf, tp, 20, ops.V3(), presolve=pre)
    u, inv = ops.unique_points(tn.view(E * P, dim))
    torch.cuda.synchronize()
os.environ.pop("MM_KT_CAP", None)
points, conn = meshgen.hex8_mesh((4, 4, 4), warp=0.02)
connC = np.ascontiguousarray(conn[:, np.argsort([0, 3, 2, 1, 4, 5, 6, 7])])
q = rng.uniform(-0.05, 1.05, (500, 3))
c8 = ops.centroid_conn(torch.from_numpy(conn).to(dev), torch.from_numpy(points).to(dev))
nn = ops.GridIndex(c8).query_idx(torch.from_numpy(q).to(dev), 20).to(torch.int64)
ops.trilinear(nn, torch.from_numpy(connC).to(dev), torch.from_numpy(points).to(dev), torch.from_numpy(q).to(dev))
torch.cuda.synchronize()
print("sanitize case done")
