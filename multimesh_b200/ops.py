"""
PyTorch custom ops over the C-ABI (include/multimesh_b200.h).

torch is plumbing here: device memory, the current CUDA stream and `torch.library` registration.
Every op takes CUDA tensors, launches the hand-written sm_100a kernels through ctypes on torch's
current stream and returns CUDA tensors.  CPU tensors are rejected -- there is no CPU path.

Inner operator boundary of the reference that these ops replace:
    KDTree(...).query(pts, k)                       -> GridIndex.query          (K1)
    inverse_transform + _check_if_inside_element... -> locate                   (K2)
    get_coefficients + np.sum(data[elem] * coeffs)  -> interp / coeffs          (K3)
    the three in one stream-ordered call             -> interpolate              (K1 -> K2 -> K3, progressive)
    lib.triLinearInterpolator / + KDTree.query      -> trilinear / trilinear_indexed (V6, order-1 nodal path)
"""
import ctypes as C
from dataclasses import dataclass
from typing import Tuple

import torch

from . import _lib
from ._lib import (FB_FAIL, FB_MAGIC, FB_MINL1, FB_SNAP, LocateParams, MultiMeshError, check,  # noqa: F401
                   load_lib)
from .gll import SUPPORTED_ORDERS

__all__ = [
    "LocateSpec", "V1", "V2", "V3", "V4", "V5", "GridIndex", "element_geometry", "locate", "interp",
    "coeffs", "gather_coeffs", "trilinear", "trilinear_indexed", "centroid_conn", "gather_nodal", "map_to_sphere_",
    "interpolate", "element_presolve", "ResidentSource", "unique_points", "scatter_back", "fluid_fixup_",
]


# ----------------------------------------------------------------------------------------------
# location-logic variants of the reference, as parameter sets of ONE kernel
# ----------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class LocateSpec:
    aabb_prefilter: bool
    tol: float
    strict: bool
    fallback: int
    snap_clip: float = 1.02
    magic_xi: Tuple[float, float, float] = (0.645, -0.5, 0.22)

    def to_c(self) -> LocateParams:
        p = LocateParams()
        p.aabb_prefilter = int(self.aabb_prefilter)
        p.strict = int(self.strict)
        p.fallback = int(self.fallback)
        p.tol = float(self.tol)
        p.snap_clip = float(self.snap_clip)
        for i in range(3):
            p.magic_xi[i] = float(self.magic_xi[i])
        return p


def V1() -> LocateSpec:
    """_check_if_inside_element (interpolator.py:1409-1473): AABB prefilter, |xi| <= 1.04,
    fallback = first AABB hit / nearest centre with the reference's magic xi."""
    return LocateSpec(True, 1.04, False, FB_MAGIC)


def V2(tolerance: float = 1.05, snap_to_nearest: bool = False) -> LocateSpec:
    """get_element_weights.check_inside (interpolator.py:1181-1233)."""
    return LocateSpec(False, tolerance, True, FB_SNAP if snap_to_nearest else FB_FAIL)


def V3() -> LocateSpec:
    """get_element_weights_layered.check_inside (interpolator.py:1271-1297)."""
    return LocateSpec(False, 1.03, True, FB_FAIL)


def V4() -> LocateSpec:
    """v2_interpolation_tools.get_element_weights (v2_interpolation_tools.py:71-164)."""
    return LocateSpec(False, 1.05, True, FB_FAIL)


def V5() -> LocateSpec:
    """cli._check_if_inside_element (scripts/cli.py:401-430)."""
    return LocateSpec(False, 1.02, False, FB_MINL1)


# ----------------------------------------------------------------------------------------------
# helpers
# ----------------------------------------------------------------------------------------------
def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(0 if t is None else t.data_ptr())


def _need_cuda(t: torch.Tensor, name: str, dtype=None, align16: bool = False) -> torch.Tensor:
    """Validates an operand; never copies.  A mis-strided or mis-aligned operand is an error, not a hidden
    `.contiguous()` / `.clone()` -- at BASELINE config 3 that would silently duplicate an 80 GB mesh."""
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name}: expected a torch.Tensor, got {type(t)}")
    if not t.is_cuda:
        raise MultiMeshError(
            f"{name}: tensor is on {t.device}; multimesh_b200 runs on CUDA only (no CPU fallback)")
    if dtype is not None and t.dtype != dtype:
        raise TypeError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name}: tensor of shape {tuple(t.shape)} with strides {t.stride()} is not contiguous; "
                         "pass a contiguous tensor (multimesh_b200 does not copy operands behind the caller's back)")
    if align16 and t.numel() and t.data_ptr() % 16 != 0:  # bulk-async copies need 16-byte aligned bases
        raise ValueError(f"{name}: data pointer {t.data_ptr():#x} is not 16-byte aligned (slice of a larger "
                         "tensor?); element blocks are fetched with bulk-async copies that need it")
    return t


def _order_dim(P: int, dim: int) -> int:
    for o in SUPPORTED_ORDERS:
        if (o + 1) ** dim == P:
            return o
    raise ValueError(f"{P} nodes per element is not (order+1)^{dim} for order in {SUPPORTED_ORDERS}")


# ----------------------------------------------------------------------------------------------
# K0
# ----------------------------------------------------------------------------------------------
@torch.library.custom_op("multimesh::element_geometry", mutates_args=())
def element_geometry(nodes: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """nodes [E,P,d] -> (centroid [E,d], aabb [E,2,d])."""
    nodes = _need_cuda(nodes, "nodes", torch.float64, align16=True)
    E, P, d = nodes.shape
    order = _order_dim(P, d)
    with torch.cuda.device(nodes.device):
        cent = torch.empty((E, d), dtype=torch.float64, device=nodes.device)
        box = torch.empty((E, 2, d), dtype=torch.float64, device=nodes.device)
        check(load_lib().mm_element_geometry(order, d, E, _ptr(nodes), _ptr(cent), _ptr(box),
                                             _stream()), "mm_element_geometry")
    return cent, box


@torch.library.custom_op("multimesh::element_presolve", mutates_args=())
def element_presolve(nodes: torch.Tensor) -> torch.Tensor:
    """nodes [E,P,d] -> presolve [E, 2d + d*d] = (first node, x(xi=0) - first node, inverse Jacobian
    at xi=0): the affine pre-solve that lets K2 start Newton one step ahead (once per source mesh)."""
    nodes = _need_cuda(nodes, "nodes", torch.float64, align16=True)
    E, P, d = nodes.shape
    order = _order_dim(P, d)
    with torch.cuda.device(nodes.device):
        pre = torch.empty((E, 2 * d + d * d), dtype=torch.float64, device=nodes.device)
        check(load_lib().mm_element_presolve(order, d, E, _ptr(nodes), _ptr(pre), _stream()),
              "mm_element_presolve")
    return pre


def map_to_sphere_(points: torch.Tensor, radius_1d: torch.Tensor, r_earth: float = 6371000.0):
    """In place; points [..., 3], radius_1d [...] (interpolator.py:1125-1144)."""
    p = _need_cuda(points, "points", torch.float64)
    r = _need_cuda(radius_1d, "radius_1d", torch.float64)
    if p.data_ptr() != points.data_ptr():
        raise ValueError("map_to_sphere_: points must be contiguous (in-place op)")
    n = r.numel()
    assert p.numel() == 3 * n
    with torch.cuda.device(p.device):
        check(load_lib().mm_map_to_sphere(n, _ptr(p), _ptr(r), float(r_earth), _stream()),
              "mm_map_to_sphere")
    return points


# ----------------------------------------------------------------------------------------------
# K1
# ----------------------------------------------------------------------------------------------
@torch.library.custom_op("multimesh::knn", mutates_args=())
def _knn_op(handle: int, pts: torch.Tensor, k: int, divisor: int) -> torch.Tensor:
    pts = _need_cuda(pts, "pts", torch.float64)
    N = pts.shape[0]
    with torch.cuda.device(pts.device):
        idx = torch.empty((N, k), dtype=torch.int32, device=pts.device)
        check(load_lib().mm_knn(C.c_void_p(handle), N, _ptr(pts), k, divisor, _ptr(idx), None,
                                _stream()), "mm_knn")
    return idx


class GridIndex:
    """GPU replacement for `pykdtree.kdtree.KDTree(data)`: exact k-NN in the canonical
    (d2, index) order.  `query` mirrors KDTree.query's (dist, idx) return."""

    def __init__(self, data: torch.Tensor):
        data = _need_cuda(data, "data", torch.float64)
        if data.dim() != 2 or data.shape[1] not in (2, 3):
            raise ValueError("GridIndex: data must be [M, 2] or [M, 3]")
        self.device = data.device
        self.n, self.dim = data.shape
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            check(load_lib().mm_index_create(C.byref(h), self.dim, self.n, _ptr(data), _stream()),
                  "mm_index_create")
        self._h = h.value
        self._sites = False

    def prepare_sites(self):
        """Builds the site table (distinct coordinates) once; used by `interpolate` in the GLL-point form
        (mm_index_prepare_sites).  Mutates the index: not to be raced with queries on other streams."""
        if not self._sites:
            with torch.cuda.device(self.device):
                check(load_lib().mm_index_prepare_sites(C.c_void_p(self._h), _stream()), "mm_index_prepare_sites")
            self._sites = True
        return self

    def info(self):
        arr = (C.c_int64 * 8)()
        cell = C.c_double()
        check(load_lib().mm_index_info(C.c_void_p(self._h), C.byref(arr), C.byref(cell)),
              "mm_index_info")
        return {"M": arr[0], "dim": arr[1], "cells": (arr[2], arr[3], arr[4]),
                "nonempty_cells": arr[5], "bytes": arr[6], "cell_size": cell.value}

    def query_idx(self, pts: torch.Tensor, k: int, divisor: int = 1) -> torch.Tensor:
        if self._h is None:
            raise MultiMeshError("GridIndex used after close()")
        if pts.dim() != 2 or pts.shape[1] != self.dim:
            raise ValueError(f"query points must be [N, {self.dim}]")
        return _knn_op(self._h, pts, int(k), int(divisor))

    def query(self, pts: torch.Tensor, k: int = 1):
        """(dist [N,k], idx [N,k]); dist = sqrt(d2)."""
        pts = _need_cuda(pts, "pts", torch.float64)
        N = pts.shape[0]
        with torch.cuda.device(self.device):
            idx = torch.empty((N, k), dtype=torch.int32, device=self.device)
            d2 = torch.empty((N, k), dtype=torch.float64, device=self.device)
            check(load_lib().mm_knn(C.c_void_p(self._h), N, _ptr(pts), int(k), 1, _ptr(idx),
                                    _ptr(d2), _stream()), "mm_knn")
        return torch.sqrt(d2), idx

    def close(self):
        if getattr(self, "_h", None):
            load_lib().mm_index_destroy(C.c_void_p(self._h))
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ----------------------------------------------------------------------------------------------
# K2
# ----------------------------------------------------------------------------------------------
@torch.library.custom_op("multimesh::locate", mutates_args=())
def _locate_op(nodes: torch.Tensor, centroid: torch.Tensor, aabb: torch.Tensor, presolve: torch.Tensor,
               pts: torch.Tensor, cands: torch.Tensor, aabb_prefilter: bool, tol: float, strict: bool, fallback: int,
               snap_clip: float, m0: float, m1: float, m2: float
               ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    nodes = _need_cuda(nodes, "nodes", torch.float64, align16=True)
    pts = _need_cuda(pts, "pts", torch.float64)
    cands = _need_cuda(cands, "cands", torch.int32)
    centroid = _need_cuda(centroid, "centroid", torch.float64)
    aabb = _need_cuda(aabb, "aabb", torch.float64)
    pre = _need_cuda(presolve, "presolve", torch.float64) if presolve.numel() else None
    E, P, d = nodes.shape
    order = _order_dim(P, d)
    N, k = cands.shape
    if pts.shape != (N, d):
        raise ValueError(f"pts must be [{N}, {d}], got {tuple(pts.shape)}")
    prm = LocateSpec(aabb_prefilter, tol, strict, fallback, snap_clip, (m0, m1, m2)).to_c()
    dev = nodes.device
    with torch.cuda.device(dev):
        elem = torch.empty((N,), dtype=torch.int32, device=dev)
        xi = torch.empty((N, d), dtype=torch.float64, device=dev)
        status = torch.empty((N,), dtype=torch.uint8, device=dev)
        nfail = torch.zeros((1,), dtype=torch.int64, device=dev)
        check(load_lib().mm_locate(order, d, E, _ptr(nodes), _ptr(centroid), _ptr(aabb), _ptr(pre), N,
                                   _ptr(pts), k, _ptr(cands), C.byref(prm), _ptr(elem), _ptr(xi),
                                   _ptr(status), _ptr(nfail), _stream()), "mm_locate")
    return elem, xi, status, nfail


def _no_tensor(like):
    return torch.empty((0,), dtype=torch.float64, device=like.device)


def locate(nodes, centroid, aabb, pts, cands, spec: LocateSpec, presolve=None):
    """-> (elem [N] i32, xi [N,d] f64, status [N] u8, num_failed [1] i64 on device).
    `presolve` (element_presolve) makes Newton start from the affine pre-solve instead of xi = 0."""
    return _locate_op(nodes, centroid, aabb, _no_tensor(nodes) if presolve is None else presolve, pts, cands,
                      spec.aabb_prefilter, spec.tol,
                      spec.strict, spec.fallback, spec.snap_clip, *spec.magic_xi)


# ----------------------------------------------------------------------------------------------
# K3
# ----------------------------------------------------------------------------------------------
@torch.library.custom_op("multimesh::interp", mutates_args=())
def interp(fields: torch.Tensor, elem: torch.Tensor, xi: torch.Tensor) -> torch.Tensor:
    """fields [E,F,P], elem [N] i32, xi [N,d] -> out [N,F]."""
    fields = _need_cuda(fields, "fields", torch.float64, align16=True)
    elem = _need_cuda(elem, "elem", torch.int32)
    xi = _need_cuda(xi, "xi", torch.float64)
    E, F, P = fields.shape
    N, d = xi.shape
    order = _order_dim(P, d)
    with torch.cuda.device(fields.device):
        out = torch.empty((N, F), dtype=torch.float64, device=fields.device)
        check(load_lib().mm_interp(order, d, E, F, _ptr(fields), N, _ptr(elem), _ptr(xi),
                                   _ptr(out), _stream()), "mm_interp")
    return out


def interp_perm(fields: torch.Tensor, elem: torch.Tensor, xi: torch.Tensor, perm) -> torch.Tensor:
    """K3, coherent variant (mm_interp_perm): row n of the result goes to out[perm[n]] (perm may be
    None).  Meant for spatially sorted points; correct for any order."""
    fields = _need_cuda(fields, "fields", torch.float64, align16=True)
    elem = _need_cuda(elem, "elem", torch.int32)
    xi = _need_cuda(xi, "xi", torch.float64)
    if perm is not None:
        perm = _need_cuda(perm, "perm", torch.int32)
    E, F, P = fields.shape
    N, d = xi.shape
    order = _order_dim(P, d)
    with torch.cuda.device(fields.device):
        out = torch.empty((N, F), dtype=torch.float64, device=fields.device)
        check(load_lib().mm_interp_perm(order, d, E, F, _ptr(fields), N, _ptr(elem), _ptr(xi),
                                        _ptr(perm), _ptr(out), _stream()), "mm_interp_perm")
    return out


@torch.library.custom_op("multimesh::coeffs", mutates_args=())
def coeffs(elem: torch.Tensor, xi: torch.Tensor, order: int) -> torch.Tensor:
    """-> coeffs [N,P]; zero rows where elem < 0."""
    elem = _need_cuda(elem, "elem", torch.int32)
    xi = _need_cuda(xi, "xi", torch.float64)
    N, d = xi.shape
    P = (order + 1) ** d
    with torch.cuda.device(xi.device):
        out = torch.empty((N, P), dtype=torch.float64, device=xi.device)
        check(load_lib().mm_coeffs(order, d, N, _ptr(elem), _ptr(xi), _ptr(out), _stream()),
              "mm_coeffs")
    return out


def gather_coeffs(fields: torch.Tensor, elem: torch.Tensor, coeff: torch.Tensor) -> torch.Tensor:
    """Cached-matrix gather: out[n,f] = sum_a fields[elem_n,f,a] * coeff[n,a]."""
    fields = _need_cuda(fields, "fields", torch.float64, align16=True)
    elem = _need_cuda(elem, "elem", torch.int32)
    coeff = _need_cuda(coeff, "coeffs", torch.float64)
    E, F, P = fields.shape
    N = elem.shape[0]
    assert coeff.shape == (N, P)
    with torch.cuda.device(fields.device):
        out = torch.empty((N, F), dtype=torch.float64, device=fields.device)
        check(load_lib().mm_gather_coeffs(P, E, F, _ptr(fields), N, _ptr(elem), _ptr(coeff),
                                          _ptr(out), _stream()), "mm_gather_coeffs")
    return out


# ----------------------------------------------------------------------------------------------
# order-1 nodal (Exodus) path
# ----------------------------------------------------------------------------------------------
def trilinear(nearest: torch.Tensor, connectivity: torch.Tensor, nodes: torch.Tensor,
              points: torch.Tensor):
    """Device twin of lib.triLinearInterpolator.  nearest [N,k] i64, connectivity [E,8] i64 (C
    vertex order), nodes [Np,3], points [N,3] -> (num_failed [1] i64, enclosing [N,8] i64,
    weights [N,8]); failed points keep zero weights / zero node ids."""
    nearest = _need_cuda(nearest, "nearest", torch.int64)
    connectivity = _need_cuda(connectivity, "connectivity", torch.int64)
    nodes = _need_cuda(nodes, "nodes", torch.float64, align16=True)
    points = _need_cuda(points, "points", torch.float64)
    N, k = nearest.shape
    dev = points.device
    with torch.cuda.device(dev):
        enc = torch.zeros((N, 8), dtype=torch.int64, device=dev)
        w = torch.zeros((N, 8), dtype=torch.float64, device=dev)
        nfail = torch.zeros((1,), dtype=torch.int64, device=dev)
        check(load_lib().mm_trilinear(k, N, _ptr(nearest), _ptr(connectivity), _ptr(enc),
                                      _ptr(nodes), _ptr(w), _ptr(points), _ptr(nfail), _stream()),
              "mm_trilinear")
    return nfail, enc, w


def trilinear_indexed(index: "GridIndex", connectivity: torch.Tensor, nodes: torch.Tensor, points: torch.Tensor,
                      k: int = 20):
    """`index.query_idx(points, k)` + `trilinear` as one stream-ordered call with the progressive search
    (mm_trilinear_indexed): same (num_failed [1] i64, enclosing [N,8] i64, weights [N,8]), bit for bit, without the
    [N,k] candidate array.  index: GridIndex over the HEX8 centroids (`centroid_conn`); connectivity [E,8] i64 in the C
    routine's vertex order."""
    connectivity = _need_cuda(connectivity, "connectivity", torch.int64)
    nodes = _need_cuda(nodes, "nodes", torch.float64, align16=True)
    points = _need_cuda(points, "points", torch.float64)
    N = points.shape[0]
    E = connectivity.shape[0]
    dev = points.device
    lib = load_lib()
    with torch.cuda.device(dev):
        enc = torch.zeros((N, 8), dtype=torch.int64, device=dev)
        w = torch.zeros((N, 8), dtype=torch.float64, device=dev)
        nfail = torch.zeros((1,), dtype=torch.int64, device=dev)
        nbytes = lib.mm_trilinear_indexed_workspace_bytes(index._h, N, int(k))
        ws = torch.empty((max(int(nbytes), 256),), dtype=torch.uint8, device=dev)
        check(lib.mm_trilinear_indexed(index._h, E, _ptr(connectivity), _ptr(nodes), N, _ptr(points), int(k),
                                       _ptr(enc), _ptr(w), _ptr(nfail), _ptr(ws), ws.numel(), _stream()),
              "mm_trilinear_indexed")
    return nfail, enc, w


def centroid_conn(connectivity: torch.Tensor, points: torch.Tensor) -> torch.Tensor:
    connectivity = _need_cuda(connectivity, "connectivity", torch.int64)
    points = _need_cuda(points, "points", torch.float64)
    E, npe = connectivity.shape
    with torch.cuda.device(points.device):
        out = torch.empty((E, points.shape[1]), dtype=torch.float64, device=points.device)
        check(load_lib().mm_centroid_conn(points.shape[1], E, npe, _ptr(connectivity),
                                          _ptr(points), _ptr(out), _stream()), "mm_centroid_conn")
    return out


def gather_nodal(param: torch.Tensor, enclosing: torch.Tensor, weights: torch.Tensor) -> torch.Tensor:
    """param [F,Np], enclosing [N,8] i64, weights [N,8] -> values [F,N]."""
    param = _need_cuda(param, "param", torch.float64)
    enclosing = _need_cuda(enclosing, "enclosing", torch.int64)
    weights = _need_cuda(weights, "weights", torch.float64)
    F, npm = param.shape
    N = enclosing.shape[0]
    with torch.cuda.device(param.device):
        out = torch.empty((F, N), dtype=torch.float64, device=param.device)
        check(load_lib().mm_gather_nodal(F, npm, _ptr(param), N, _ptr(enclosing), _ptr(weights),
                                         _ptr(out), _stream()), "mm_gather_nodal")
    return out


# ----------------------------------------------------------------------------------------------
# fused pipeline K1 -> K2 -> K3 (mm_interpolate): spatially sorted points, progressive search
# ----------------------------------------------------------------------------------------------
@torch.library.custom_op("multimesh::interpolate", mutates_args=("out_buf",))
def _interpolate_op(handle: int, divisor: int, nodes: torch.Tensor, centroid: torch.Tensor,
                    aabb: torch.Tensor, presolve: torch.Tensor, fields: torch.Tensor, pts: torch.Tensor, k: int,
                    aabb_prefilter: bool, tol: float, strict: bool, fallback: int, snap_clip: float,
                    m0: float, m1: float, m2: float, want_location: bool, out_buf: torch.Tensor
                    ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    nodes = _need_cuda(nodes, "nodes", torch.float64, align16=True)
    centroid = _need_cuda(centroid, "centroid", torch.float64)
    aabb = _need_cuda(aabb, "aabb", torch.float64)
    pts = _need_cuda(pts, "pts", torch.float64)
    pre = _need_cuda(presolve, "presolve", torch.float64) if presolve.numel() else None
    E, P, d = nodes.shape
    order = _order_dim(P, d)
    N = pts.shape[0]
    have_fields = fields.numel() > 0
    if have_fields:
        fields = _need_cuda(fields, "fields", torch.float64, align16=True)
        F = fields.shape[1]
        assert fields.shape[0] == E and fields.shape[2] == P
    else:
        F = 1
    prm = LocateSpec(aabb_prefilter, tol, strict, fallback, snap_clip, (m0, m1, m2)).to_c()
    dev = nodes.device
    lib = load_lib()
    with torch.cuda.device(dev):
        nbytes = lib.mm_interpolate_workspace_bytes(C.c_void_p(handle), d, N, k)
        ws = torch.empty((max(int(nbytes), 256),), dtype=torch.uint8, device=dev)
        if have_fields and out_buf.numel():
            # the caller's buffer (e.g. this rank's rows of the gather buffer): K3 writes straight into it
            out = _need_cuda(out_buf, "out", torch.float64)
            if tuple(out.shape) != (N, F):
                raise ValueError(f"out must be [{N}, {F}], got {tuple(out.shape)}")
        else:
            out = torch.empty((N, F) if have_fields else (0, 0), dtype=torch.float64, device=dev)
        nfail = torch.zeros((1,), dtype=torch.int64, device=dev)
        if want_location:
            elem = torch.empty((N,), dtype=torch.int32, device=dev)
            xi = torch.empty((N, d), dtype=torch.float64, device=dev)
            status = torch.empty((N,), dtype=torch.uint8, device=dev)
        else:
            elem = torch.empty((0,), dtype=torch.int32, device=dev)
            xi = torch.empty((0, d), dtype=torch.float64, device=dev)
            status = torch.empty((0,), dtype=torch.uint8, device=dev)
        check(lib.mm_interpolate(C.c_void_p(handle), divisor, order, d, E, _ptr(nodes), _ptr(centroid),
                                 _ptr(aabb), _ptr(pre), F, _ptr(fields) if have_fields else None, N, _ptr(pts), k,
                                 C.byref(prm), _ptr(out) if have_fields else None,
                                 _ptr(elem) if want_location else None, _ptr(xi) if want_location else None,
                                 _ptr(status) if want_location else None, _ptr(nfail), _ptr(ws),
                                 ws.numel(), _stream()), "mm_interpolate")
    if have_fields and out_buf.numel():
        out = torch.empty((0, 0), dtype=torch.float64, device=dev)  # results are in out_buf (an op may not return an input)
    return out, elem, xi, status, nfail


def interpolate(index: "GridIndex", divisor: int, nodes, centroid, aabb, fields, pts, k: int,
                spec: LocateSpec, want_location: bool = True, presolve=None, out=None):
    """Fused k-NN -> locate -> gather.  `fields` may be None (locate only).
    -> (out [N,F], elem [N], xi [N,d], status [N], num_failed [1]); identical to running
    GridIndex.query_idx, locate and interp one after the other.  Stream-ordered (no host sync).
    `out`: optional [N,F] buffer the gather writes into (e.g. this rank's slice of a gather buffer)."""
    if fields is None:
        fields = torch.empty((0, 0, 0), dtype=torch.float64, device=nodes.device)
    if divisor > 1:
        index.prepare_sites()
    res = _interpolate_op(index._h, int(divisor), nodes, centroid, aabb,
                          _no_tensor(nodes) if presolve is None else presolve, fields, pts, int(k),
                          spec.aabb_prefilter, spec.tol, spec.strict, spec.fallback, spec.snap_clip,
                          *spec.magic_xi, bool(want_location), _no_tensor(nodes) if out is None else out)
    if out is not None and fields.numel():
        return (out,) + tuple(res[1:])
    return res


# ----------------------------------------------------------------------------------------------
# K4: target-side de-duplication and write-back
# ----------------------------------------------------------------------------------------------
def unique_points(pts: torch.Tensor):
    """Device twin of np.unique(pts, axis=0, return_inverse=True) (utils.get_unique_points, utils.py:465-515):
    pts [N,d] -> (unique [N_u,d] in lexicographic order, inverse [N] int32)."""
    pts = _need_cuda(pts, "pts", torch.float64)
    N, d = pts.shape
    with torch.cuda.device(pts.device):
        uniq = torch.empty((N, d), dtype=torch.float64, device=pts.device)
        inv = torch.empty((N,), dtype=torch.int32, device=pts.device)
        n = C.c_int64(0)
        check(load_lib().mm_unique_points(d, N, _ptr(pts), C.byref(n), _ptr(uniq), _ptr(inv), _stream()),
              "mm_unique_points")
    return uniq[: n.value], inv


def scatter_back(values: torch.Tensor, inverse, E: int, P: int) -> torch.Tensor:
    """values [N_u,F] (+ inverse [E*P] int32, or None for the identity) -> [E,F,P]: the
    values[recon].reshape(E, P, F).swapaxes(1, 2) of the reference's drivers (interpolator.py:822-826)."""
    values = _need_cuda(values, "values", torch.float64)
    F = values.shape[1]
    if inverse is not None:
        inverse = _need_cuda(inverse, "inverse", torch.int32)
        assert inverse.numel() == E * P
    else:
        assert values.shape[0] == E * P
    with torch.cuda.device(values.device):
        out = torch.empty((E, F, P), dtype=torch.float64, device=values.device)
        check(load_lib().mm_scatter_back(E, P, F, _ptr(values), _ptr(inverse), _ptr(out), _stream()), "mm_scatter_back")
    return out


def fluid_fixup_(values: torch.Tensor, old_values: torch.Tensor, fluid: torch.Tensor, vs_index: int) -> torch.Tensor:
    """In place on values [E,F,P]: fluid elements keep old_values, solid elements that picked up VS == 0 are
    restored (interpolator.py:829-841).  fluid: uint8 / bool [E]."""
    values = _need_cuda(values, "values", torch.float64)
    old_values = _need_cuda(old_values, "old_values", torch.float64)
    fluid = _need_cuda(fluid.to(torch.uint8), "fluid", torch.uint8)
    E, F, P = values.shape
    assert old_values.shape == values.shape and fluid.numel() == E
    with torch.cuda.device(values.device):
        check(load_lib().mm_fluid_fixup(E, P, F, _ptr(values), _ptr(old_values), _ptr(fluid), int(vs_index), _stream()),
              "mm_fluid_fixup")
    return values


class ResidentSource:
    """A source mesh resident in HBM (mm_source_*): nodes, fields, geometry, pre-solve and the spatial index
    are uploaded / built once; `interpolate_host` then moves only target points in and values out, chunked
    over three streams so that both PCIe directions and the kernels overlap.  Host (numpy / pinned torch CPU)
    arrays in and out -- the host-buffer twin of `interpolate`."""

    def __init__(self, nodes, fields=None, form: str = "centroid", device=None):
        import numpy as np

        self._np = np
        self._h = None
        nodes = self._host(nodes)
        E, P, d = nodes.shape
        order = _order_dim(P, d)
        F = 0
        if fields is not None:
            fields = self._host(fields)
            assert fields.shape[0] == E and fields.shape[2] == P
            F = fields.shape[1]
        self.E, self.P, self.dim, self.F, self.order = E, P, d, F, order
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            check(load_lib().mm_source_create_host(C.byref(h), order, d, E, self._p(nodes), F,
                                                   self._p(fields) if F else None, 1 if form == "gll" else 0),
                  "mm_source_create_host")
        self._h = h.value

    def _host(self, a):
        if isinstance(a, torch.Tensor):
            if a.is_cuda or a.dtype != torch.float64 or not a.is_contiguous():
                raise ValueError("ResidentSource: host arrays must be contiguous float64 CPU tensors / numpy arrays")
            return a
        return self._np.ascontiguousarray(a, dtype=self._np.float64)

    @staticmethod
    def _p(a):
        if a is None:
            return None
        return C.c_void_p(a.data_ptr() if isinstance(a, torch.Tensor) else a.ctypes.data)

    def set_fields(self, fields):
        fields = self._host(fields)
        assert fields.shape[0] == self.E and fields.shape[2] == self.P
        check(load_lib().mm_source_set_fields_host(C.c_void_p(self._h), fields.shape[1], self._p(fields)),
              "mm_source_set_fields_host")
        self.F = fields.shape[1]

    def info(self):
        arr = (C.c_int64 * 8)()
        check(load_lib().mm_source_info(C.c_void_p(self._h), C.byref(arr)), "mm_source_info")
        return {"E": arr[0], "P": arr[1], "dim": arr[2], "F": arr[3], "order": arr[4], "gll_points_form": arr[5],
                "resident_bytes": arr[6], "device": arr[7]}

    def interpolate_host(self, pts, k: int, spec: LocateSpec, values=None, want_location: bool = False):
        """pts [N,d] host -> (values [N,F], elem [N] | None, xi [N,d] | None, num_failed).  `values` may be a
        preallocated (ideally pinned) host buffer."""
        np = self._np
        pts = self._host(pts)
        N = pts.shape[0]
        if values is None:
            values = np.empty((N, self.F), dtype=np.float64)
        elem = np.empty((N,), dtype=np.int32) if want_location else None
        xi = np.empty((N, self.dim), dtype=np.float64) if want_location else None
        prm = spec.to_c()
        nf = C.c_int64(0)
        check(load_lib().mm_source_interpolate_host(C.c_void_p(self._h), N, self._p(pts), int(k), C.byref(prm),
                                                    self._p(values), self._p(elem), self._p(xi), C.byref(nf)),
              "mm_source_interpolate_host")
        return values, elem, xi, int(nf.value)

    def close(self):
        if getattr(self, "_h", None):
            load_lib().mm_source_destroy(C.c_void_p(self._h))
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
