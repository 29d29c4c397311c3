"""
KDTree -- GPU stand-in for `pykdtree.kdtree.KDTree` with the call shape the reference uses:
    tree = KDTree(data[M, d]);  dist, idx = tree.query(pts[N, d], k)
Backed by the counting-sorted uniform grid of mm_index.cu; neighbours come back in the canonical
(d2, index) order (numpy in, numpy out; device tensors are accepted and returned as given).
"""
import numpy as np
import torch

from . import ops


def _device(device=None):
    if device is not None:
        return torch.device(device)
    if not torch.cuda.is_available():
        raise ops.MultiMeshError("multimesh_b200 needs a CUDA device (no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


class KDTree(object):
    def __init__(self, data, device=None):
        self.device = data.device if isinstance(data, torch.Tensor) and data.is_cuda else _device(device)
        t = data if isinstance(data, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(data, dtype=np.float64))
        self.data = data
        self._t = t.to(self.device, dtype=torch.float64)
        self.n, self.m = self._t.shape
        self.index = ops.GridIndex(self._t)

    def query(self, pts, k=1, divisor=1, return_distance=True):
        as_numpy = not isinstance(pts, torch.Tensor)
        t = torch.from_numpy(np.ascontiguousarray(pts, dtype=np.float64)) if as_numpy else pts
        t = t.to(self.device, dtype=torch.float64)
        if return_distance:
            dist, idx = self.index.query(t, k)
            if divisor != 1:
                idx = torch.div(idx, divisor, rounding_mode="floor")
        else:
            dist, idx = None, self.index.query_idx(t, k, divisor)
        if as_numpy:
            idx = idx.cpu().numpy().astype(np.int64)
            dist = None if dist is None else dist.cpu().numpy()
            if k == 1:
                idx = idx[:, 0]
                dist = None if dist is None else dist[:, 0]
        return dist, idx
