"""
ctypes bindings for the C oracle (oracle/mm_oracle.c) and for the compiled reference
(oracle/_ref).  TEST INFRASTRUCTURE ONLY -- never imported by multimesh_b200/.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

_f64 = np.ctypeslib.ndpointer(dtype=np.float64, flags=["C_CONTIGUOUS"])
_i32 = np.ctypeslib.ndpointer(dtype=np.int32, flags=["C_CONTIGUOUS"])
_i64 = np.ctypeslib.ndpointer(dtype=np.int64, flags=["C_CONTIGUOUS"])
_u8 = np.ctypeslib.ndpointer(dtype=np.uint8, flags=["C_CONTIGUOUS"])

FB_FAIL, FB_MAGIC, FB_SNAP, FB_MINL1 = 0, 1, 2, 3
(ST_ACCEPTED, ST_FB_INSIDE_MAGIC, ST_FB_NEAR_OK, ST_FB_NEAR_MAGIC, ST_FB_NAN_MAGIC,
 ST_SNAPPED, ST_FAILED, ST_MINL1, ST_SNAP_NONE) = range(9)


class LocateParams(C.Structure):
    _fields_ = [
        ("aabb_prefilter", C.c_int32),
        ("strict", C.c_int32),
        ("fallback", C.c_int32),
        ("reserved", C.c_int32),
        ("tol", C.c_double),
        ("snap_clip", C.c_double),
        ("magic_xi", C.c_double * 3),
    ]


def params(aabb_prefilter, tol, strict, fallback, snap_clip=1.02, magic_xi=(0.645, -0.5, 0.22)):
    p = LocateParams()
    p.aabb_prefilter = int(aabb_prefilter)
    p.strict = int(strict)
    p.fallback = int(fallback)
    p.tol = float(tol)
    p.snap_clip = float(snap_clip)
    for i in range(3):
        p.magic_xi[i] = float(magic_xi[i])
    return p


# The reference's location-logic variants
def V1():  # _check_if_inside_element, interpolator.py:1409-1473
    return params(True, 1.04, False, FB_MAGIC)


def V2(tolerance=1.05, snap_to_nearest=False):  # get_element_weights, :1181-1233
    return params(False, tolerance, True, FB_SNAP if snap_to_nearest else FB_FAIL)


def V3():  # get_element_weights_layered, :1271-1297
    return params(False, 1.03, True, FB_FAIL)


def V4():  # v2_interpolation_tools.py:71-164
    return params(False, 1.05, True, FB_FAIL)


def V5():  # scripts/cli.py:401-430
    return params(False, 1.02, False, FB_MINL1)


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    so = _build.build_oracle()
    L = C.CDLL(so)
    L.mmo_gll_nodes.restype = C.c_int
    L.mmo_gll_nodes.argtypes = [C.c_int, _f64]
    L.mmo_lagrange.restype = C.c_int
    L.mmo_lagrange.argtypes = [C.c_int, C.c_double, _f64, _f64]
    L.mmo_weights.restype = C.c_int
    L.mmo_weights.argtypes = [C.c_int, C.c_int, _f64, _f64]
    L.mmo_coeffs.restype = C.c_int
    L.mmo_coeffs.argtypes = [C.c_int, C.c_int, C.c_longlong, C.c_void_p, _f64, _f64]
    L.mmo_forward_map.restype = C.c_int
    L.mmo_forward_map.argtypes = [C.c_int, C.c_int, _f64, _f64, _f64]
    L.mmo_inverse_map.restype = C.c_int
    L.mmo_inverse_map.argtypes = [C.c_int, C.c_int, _f64, _f64, _f64, C.POINTER(C.c_int)]
    L.mmo_centroids.restype = None
    L.mmo_centroids.argtypes = [C.c_longlong, C.c_int, C.c_int, _f64, _f64]
    L.mmo_aabb.restype = None
    L.mmo_aabb.argtypes = [C.c_longlong, C.c_int, C.c_int, _f64, _f64]
    L.mmo_centroid.restype = None
    L.mmo_centroid.argtypes = [C.c_longlong, C.c_longlong, C.c_longlong, _i64, _f64, _f64]
    L.mmo_knn_bruteforce.restype = None
    L.mmo_knn_bruteforce.argtypes = [C.c_longlong, C.c_int, _f64, C.c_longlong, _f64, C.c_int,
                                     _i32, C.c_void_p]
    L.mmo_locate.restype = C.c_longlong
    L.mmo_locate.argtypes = [C.c_int, C.c_int, C.c_longlong, _f64, _f64, _f64, C.c_void_p, C.c_longlong,
                             _f64, C.c_int, _i32, C.POINTER(LocateParams), _i32, _f64, _u8]
    L.mmo_presolve.restype = None
    L.mmo_presolve.argtypes = [C.c_int, C.c_int, C.c_longlong, _f64, _f64]
    L.mmo_inverse_map_pre.restype = C.c_int
    L.mmo_inverse_map_pre.argtypes = [C.c_int, C.c_int, _f64, _f64, _f64, _f64, C.POINTER(C.c_int)]
    L.mmo_interp.restype = C.c_int
    L.mmo_interp.argtypes = [C.c_int, C.c_int, C.c_longlong, C.c_int, _f64, C.c_longlong, _i32,
                             _f64, _f64]
    L.mmo_trilinear_interpolator.restype = C.c_longlong
    L.mmo_trilinear_interpolator.argtypes = [C.c_longlong, C.c_longlong, _i64, _i64, _i64, _f64,
                                             _f64, _f64]
    L.mmo_hex8_weights.restype = None
    L.mmo_hex8_weights.argtypes = [_f64, _f64]
    L.mmo_hex8_inverse.restype = C.c_int
    L.mmo_hex8_inverse.argtypes = [_f64, _f64, _f64]
    L.mmo_num_threads.restype = C.c_int
    L.mmo_set_num_threads.restype = None
    L.mmo_set_num_threads.argtypes = [C.c_int]
    _lib = L
    return L


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def num_threads():
    return int(lib().mmo_num_threads())


def set_num_threads(n):
    """torchrun exports OMP_NUM_THREADS=1; the CPU baseline wants all host threads."""
    lib().mmo_set_num_threads(int(n))


def gll_nodes(order):
    z = np.zeros(5)
    m = lib().mmo_gll_nodes(order, z)
    return z[:m].copy()


def lagrange(order, x):
    L = np.zeros(5)
    dL = np.zeros(5)
    m = lib().mmo_lagrange(order, float(x), L, dL)
    return L[:m].copy(), dL[:m].copy()


def weights(order, dim, xi):
    m = order + 1
    w = np.zeros(m ** dim)
    assert lib().mmo_weights(order, dim, _c(xi, np.float64), w) == 0
    return w


def coeffs(order, dim, elem, xi):
    xi = _c(xi, np.float64).reshape(-1, dim)
    N = xi.shape[0]
    out = np.zeros((N, (order + 1) ** dim))
    e = None if elem is None else _c(elem, np.int32)
    ep = None if e is None else e.ctypes.data_as(C.c_void_p)
    assert lib().mmo_coeffs(order, dim, N, ep, xi, out) == 0
    return out


def forward_map(order, dim, nodes, xi):
    x = np.zeros(dim)
    assert lib().mmo_forward_map(order, dim, _c(nodes, np.float64), _c(xi, np.float64), x) == 0
    return x


def inverse_map(order, dim, nodes, p):
    xi = np.zeros(dim)
    it = C.c_int(0)
    ok = lib().mmo_inverse_map(order, dim, _c(nodes, np.float64), _c(p, np.float64), xi,
                               C.byref(it))
    return bool(ok == 1), xi, it.value


def inverse_map_presolved(order, dim, nodes, p):
    """One element, Newton started from that element's affine pre-solve (the start mmo_locate uses)."""
    nodes = _c(nodes, np.float64).reshape(1, -1, dim)
    pre = presolve(nodes)
    xi = np.zeros(dim)
    it = C.c_int(0)
    ok = lib().mmo_inverse_map_pre(order, dim, nodes, _c(p, np.float64), pre, xi, C.byref(it))
    return bool(ok == 1), xi


def centroids(nodes):
    nodes = _c(nodes, np.float64)
    E, P, d = nodes.shape
    out = np.zeros((E, d))
    lib().mmo_centroids(E, P, d, nodes, out)
    return out


def aabb(nodes):
    nodes = _c(nodes, np.float64)
    E, P, d = nodes.shape
    out = np.zeros((E, 2, d))
    lib().mmo_aabb(E, P, d, nodes, out)
    return out


def centroid_conn(conn, points):
    conn = _c(conn, np.int64)
    points = _c(points, np.float64)
    E, npe = conn.shape
    out = np.zeros((E, points.shape[1]))
    lib().mmo_centroid(points.shape[1], E, npe, conn, points, out)
    return out


def knn_bruteforce(data, pts, k, return_d2=False):
    data = _c(data, np.float64)
    pts = _c(pts, np.float64)
    M, d = data.shape
    N = pts.shape[0]
    idx = np.zeros((N, k), dtype=np.int32)
    d2 = np.zeros((N, k)) if return_d2 else None
    lib().mmo_knn_bruteforce(M, d, data, N, pts, k, idx,
                             None if d2 is None else d2.ctypes.data_as(C.c_void_p))
    return (idx, d2) if return_d2 else idx


def knn_ckdtree_canonical(data, pts, k, pad=64, workers=-1):
    """k-NN via scipy.spatial.cKDTree (the KD-tree the reference's cli.py:66 uses; stands in for
    the absent pykdtree), re-ordered into the canonical (d2, index) total order.  Points whose
    k-th distance is tied beyond the over-query window fall back to brute force."""
    from scipy.spatial import cKDTree

    data = _c(data, np.float64)
    pts = _c(pts, np.float64)
    M = data.shape[0]
    kq = min(M, k + pad)
    tree = cKDTree(data)
    _, nn = tree.query(pts, k=kq, workers=workers)
    nn = nn.reshape(pts.shape[0], kq)
    diff = pts[:, None, :] - data[nn]
    d2 = diff[..., 0] * diff[..., 0] + diff[..., 1] * diff[..., 1]
    if data.shape[1] == 3:
        d2 = d2 + diff[..., 2] * diff[..., 2]
    order = np.lexsort((nn, d2), axis=1)
    nn_s = np.take_along_axis(nn, order, axis=1)
    d2_s = np.take_along_axis(d2, order, axis=1)
    out = np.full((pts.shape[0], k), -1, dtype=np.int32)
    kk = min(k, kq)
    out[:, :kk] = nn_s[:, :kk]
    if kq > k:
        # window must strictly exceed the k-th distance, else ties may be cut arbitrarily
        unsafe = np.nonzero(~(d2_s[:, kq - 1] > d2_s[:, k - 1]))[0]
        if unsafe.size:
            out[unsafe] = knn_bruteforce(data, pts[unsafe], k)
    return out


def presolve(nodes):
    """Affine pre-solve per element: [E, 2d + d*d] = (ref node, x(0) - ref, Jinv(0)); K0 output."""
    nodes = _c(nodes, np.float64)
    E, P, d = nodes.shape
    order = round(P ** (1.0 / d)) - 1
    out = np.zeros((E, 2 * d + d * d))
    lib().mmo_presolve(order, d, E, nodes, out)
    return out


def locate(order, dim, nodes, pts, cands, prm, cent=None, box=None, pre=True):
    """pre=True: start Newton from the affine pre-solve (what the drivers / pipeline do);
    pre=False: start from xi = 0; or pass a presolve array."""
    nodes = _c(nodes, np.float64)
    pts = _c(pts, np.float64).reshape(-1, dim)
    cands = _c(cands, np.int32)
    N, k = cands.shape
    E = nodes.shape[0]
    if cent is None:
        cent = centroids(nodes)
    if box is None:
        box = aabb(nodes)
    elem = np.zeros(N, dtype=np.int32)
    xi = np.zeros((N, dim))
    status = np.zeros(N, dtype=np.uint8)
    if pre is True:
        pre = presolve(nodes)
    prep = None if pre is False or pre is None else _c(pre, np.float64).ctypes.data_as(C.c_void_p)
    nfailed = lib().mmo_locate(order, dim, E, nodes, _c(cent, np.float64), _c(box, np.float64), prep, N,
                               pts, k, cands, C.byref(prm), elem, xi, status)
    assert nfailed >= 0
    return elem, xi, status, int(nfailed)


def interp(order, dim, fields, elem, xi):
    fields = _c(fields, np.float64)
    E, F, P = fields.shape
    assert P == (order + 1) ** dim
    elem = _c(elem, np.int32)
    xi = _c(xi, np.float64).reshape(-1, dim)
    out = np.zeros((elem.shape[0], F))
    assert lib().mmo_interp(order, dim, E, F, fields, elem.shape[0], elem, xi, out) == 0
    return out


def trilinear_interpolator(k, nearest, conn, nodes, points):
    """C-compat order-1 path; same argument meaning as the reference's triLinearInterpolator."""
    nearest = _c(nearest, np.int64)
    conn = _c(conn, np.int64)
    nodes = _c(nodes, np.float64)
    points = _c(points, np.float64)
    N = points.shape[0]
    enclosing = np.zeros((N, 8), dtype=np.int64)
    w = np.zeros((N, 8))
    nfailed = lib().mmo_trilinear_interpolator(k, N, nearest, conn, enclosing, nodes, w, points)
    return int(nfailed), enclosing, w


def hex8_weights(q):
    w = np.zeros(8)
    lib().mmo_hex8_weights(_c(q, np.float64), w)
    return w


def hex8_inverse(p, vtx):
    sol = np.zeros(3)
    ok = lib().mmo_hex8_inverse(_c(p, np.float64), _c(vtx, np.float64).reshape(24), sol)
    return bool(ok), sol


# ----------------------------------------------------------------------------------------------
# The compiled reference (oracle/_ref): multi_mesh/src/{centroid,trilinearinterpolator}.c
# ----------------------------------------------------------------------------------------------
_ref = None


def ref_lib():
    """Returns the ctypes handle of the compiled reference or None when it is unavailable.
    argtypes use long long, the types the C code actually declares (centroid.c:4-6,
    trilinearinterpolator.c:41-42); helpers.py:44-47 declares c_int, which only works by
    ABI accident for small values."""
    global _ref
    if _ref is not None:
        return _ref
    so = _build.build_ref()
    if so is None or not os.path.exists(so):
        return None
    L = C.CDLL(so)
    L.centroid.restype = None
    L.centroid.argtypes = [C.c_longlong, C.c_longlong, C.c_longlong, _i64, _f64, _f64]
    L.triLinearInterpolator.restype = C.c_longlong
    L.triLinearInterpolator.argtypes = [C.c_longlong, C.c_longlong, _i64, _i64, _i64, _f64, _f64,
                                        _f64]
    _ref = L
    return L


def ref_centroid(conn, points):
    L = ref_lib()
    conn = _c(conn, np.int64)
    points = _c(points, np.float64)
    out = np.zeros((conn.shape[0], points.shape[1]))
    L.centroid(points.shape[1], conn.shape[0], conn.shape[1], conn, points, out)
    return out


def ref_trilinear_interpolator(k, nearest, conn, nodes, points):
    L = ref_lib()
    nearest = _c(nearest, np.int64)
    conn = _c(conn, np.int64)
    nodes = _c(nodes, np.float64)
    points = _c(points, np.float64)
    N = points.shape[0]
    enclosing = np.zeros((N, 8), dtype=np.int64)
    w = np.zeros((N, 8))
    # the reference printf()s "not any ..." for failed points; harmless
    nfailed = L.triLinearInterpolator(k, N, nearest, conn, enclosing, nodes, w, points)
    return int(nfailed), enclosing, w
