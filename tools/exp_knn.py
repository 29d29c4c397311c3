import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from multimesh_b200 import meshgen, ops
import bench
w = dict(bench.WORKLOADS['medium'], name='medium')
nodes_h, fields_h = bench.make_source(w); pts_h = bench.make_targets(w)
dev = torch.device('cuda:0')
nodes = torch.from_numpy(nodes_h).to(dev); pts = torch.from_numpy(pts_h).to(dev)
E, P = nodes.shape[0], nodes.shape[1]
cent, box = ops.element_geometry(nodes)
def timeit(f, n=3):
    f(); torch.cuda.synchronize(); t=time.perf_counter()
    for _ in range(n): r=f()
    torch.cuda.synchronize(); return (time.perf_counter()-t)/n*1e3, r
for form, ix, div in (('gll', ops.GridIndex(nodes.view(E*P,3)), P), ('centroid', ops.GridIndex(cent), 1)):
    print(form, ix.info())
    for k in (1, 4, 8, 12, 20):
        ms, cands = timeit(lambda: ix.query_idx(pts, k, divisor=div))
        ms2, (elem, xi, st, nf) = timeit(lambda: ops.locate(nodes, cent, box, pts, cands, ops.LocateSpec(True, 1.04, False, ops.FB_FAIL)))
        print(f"  k={k:2d} knn {ms:7.3f} ms  locate(V1,no fallback) {ms2:6.3f} ms  unresolved {(elem<0).float().mean().item():.4f}")
