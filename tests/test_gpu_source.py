"""
GPU tests of the resident-source handle (mm_source_*), the chunked host pipeline and the stream-ordered
re-run of unresolved points -- all through the C-ABI, all against the CPU oracle.
"""
import ctypes as C
import os

import numpy as np
import pytest

from multimesh_b200 import meshgen

pytestmark = pytest.mark.gpu
NAMES = ["QKAPPA", "QMU", "RHO", "VP", "VS"]


def _case(order, n, rng, npts, warp=0.03, lo=-0.08, hi=1.08):
    nodes = meshgen.box_mesh((n,) * 3, order, warp=warp)
    fields = meshgen.analytic_fields(nodes, NAMES)
    pts = np.concatenate([rng.uniform(lo, hi, (npts, 3)), nodes.reshape(-1, 3)[::5]])
    return nodes, fields, np.ascontiguousarray(pts)


def _oracle_run(oracle, order, nodes, fields, pts, k, prm, form):
    P = nodes.shape[1]
    if form == "gll":
        cands = (oracle.knn_bruteforce(nodes.reshape(-1, 3), pts, k) // P).astype(np.int32)
    else:
        cands = oracle.knn_bruteforce(oracle.centroids(nodes), pts, k)
    elem, xi, st, nf = oracle.locate(order, 3, nodes, pts, cands, prm)
    return oracle.interp(order, 3, fields, elem, xi), elem, xi, nf


@pytest.mark.parametrize("order,form", [(2, "gll"), (2, "centroid"), (4, "centroid")])
def test_resident_source_host_pipeline_matches_oracle(cuda, oracle, order, form, monkeypatch):
    """mm_source_create_host + mm_source_interpolate_host with a chunk size that cuts the points into many
    chunks (and several re-run rounds) == oracle, bit for bit; independent of the chunking."""
    from multimesh_b200 import ops

    rng = np.random.default_rng(40 + order)
    nodes, fields, pts = _case(order, 5 if order == 4 else 8, rng, 6000)
    want, w_elem, w_xi, w_nf = _oracle_run(oracle, order, nodes, fields, pts, 20, oracle.V1(), form)
    src = ops.ResidentSource(nodes, fields, form=form)
    info = src.info()
    assert info["E"] == nodes.shape[0] and info["F"] == 5 and info["resident_bytes"] >= nodes.nbytes + fields.nbytes
    results = []
    for chunk, rerun in (("1000000", None), ("777", "64"), ("2048", "100000")):
        monkeypatch.setenv("MM_HOST_CHUNK", chunk)
        if rerun:
            monkeypatch.setenv("MM_RERUN_CHUNK", rerun)
        else:
            monkeypatch.delenv("MM_RERUN_CHUNK", raising=False)
        vals, elem, xi, nf = src.interpolate_host(pts, 20, ops.V1(), want_location=True)
        assert np.array_equal(elem, w_elem) and np.array_equal(xi, w_xi) and nf == w_nf
        assert np.array_equal(vals, want)
        results.append(vals)
    # values only; new fields on the same geometry
    vals, elem, xi, nf = src.interpolate_host(pts, 20, ops.V1())
    assert elem is None and np.array_equal(vals, want)
    src.set_fields(2.0 * fields[:, :3, :])
    vals3, _, _, _ = src.interpolate_host(pts, 20, ops.V1())
    assert vals3.shape == (len(pts), 3)
    assert np.array_equal(vals3, oracle.interp(order, 3, np.ascontiguousarray(2.0 * fields[:, :3, :]), w_elem, w_xi))
    src.close()


def test_one_shot_host_entry_point_variants(cuda, oracle, monkeypatch):
    """mm_interpolate_host (upload + build + chunked pipeline + release) for V1 and V2-snap, chunked."""
    from multimesh_b200 import _lib, ops

    lib = _lib.load_lib()
    rng = np.random.default_rng(5)
    nodes, fields, pts = _case(2, 7, rng, 5000, lo=-0.15, hi=1.15)
    monkeypatch.setenv("MM_HOST_CHUNK", "1500")
    for spec, prm, form in ((ops.V1(), oracle.V1(), 1), (ops.V2(1.05, True), oracle.V2(1.05, True), 0),
                            (ops.V3(), oracle.V3(), 0)):
        want, w_elem, w_xi, w_nf = _oracle_run(oracle, 2, nodes, fields, pts, 20, prm, "gll" if form else "centroid")
        vals = np.empty((len(pts), 5))
        elem = np.empty(len(pts), dtype=np.int32)
        xi = np.empty((len(pts), 3))
        nf = C.c_int64(-1)
        c = spec.to_c()
        rc = lib.mm_interpolate_host(2, 3, nodes.shape[0], nodes.ctypes.data_as(C.c_void_p), 5,
                                     fields.ctypes.data_as(C.c_void_p), len(pts), pts.ctypes.data_as(C.c_void_p), 20,
                                     form, C.byref(c), vals.ctypes.data_as(C.c_void_p),
                                     elem.ctypes.data_as(C.c_void_p), xi.ctypes.data_as(C.c_void_p), C.byref(nf))
        _lib.check(rc, "mm_interpolate_host")
        assert np.array_equal(elem, w_elem) and np.array_equal(xi, w_xi) and np.array_equal(vals, want)
        assert nf.value == w_nf
    assert lib.mm_host_release() == 0


def test_pipeline_is_stream_ordered_and_graph_capturable(cuda, oracle, monkeypatch):
    """mm_interpolate performs no host synchronisation: it can be captured into a CUDA graph and replayed;
    the re-run of unresolved points (several rounds forced) reads its work-list length on the device."""
    import torch
    from multimesh_b200 import ops

    monkeypatch.setenv("MM_RERUN_CHUNK", "256")
    rng = np.random.default_rng(11)
    nodes, fields, pts = _case(2, 6, rng, 4000, lo=-0.2, hi=1.2)  # many outside points -> many unresolved
    want, w_elem, w_xi, w_nf = _oracle_run(oracle, 2, nodes, fields, pts, 20, oracle.V1(), "gll")
    tn, tf, tp = (torch.from_numpy(a).to(cuda) for a in (nodes, fields, pts))
    cent, box = ops.element_geometry(tn)
    pre = ops.element_presolve(tn)
    E, P, _ = nodes.shape
    index = ops.GridIndex(tn.view(E * P, 3)).prepare_sites()
    out = torch.empty((len(pts), 5), dtype=torch.float64, device=cuda)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        res = ops.interpolate(index, P, tn, cent, box, tf, tp, 20, ops.V1(), presolve=pre, out=out)  # warm-up
    s.synchronize()
    assert res[0] is out and np.array_equal(out.cpu().numpy(), want)
    g = torch.cuda.CUDAGraph()
    out.zero_()
    with torch.cuda.graph(g, stream=s):
        res = ops.interpolate(index, P, tn, cent, box, tf, tp, 20, ops.V1(), presolve=pre, out=out)
    out.zero_()
    g.replay()
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), want)
    assert np.array_equal(res[1].cpu().numpy(), w_elem) and int(res[4].item()) == w_nf


def test_size_guards(cuda):
    from multimesh_b200 import _lib

    lib = _lib.load_lib()
    prm = _lib.LocateParams()
    rc = lib.mm_interpolate(None, 1, 2, 3, 1, None, None, None, None, 1, None, 2 ** 31, None, 20, C.byref(prm), None,
                            None, None, None, None, None, 0, None)
    assert rc == -1
    import torch
    idx = C.c_void_p()
    x = torch.zeros((4, 3), dtype=torch.float64, device=cuda)
    _lib.check(lib.mm_index_create(C.byref(idx), 3, 4, C.c_void_p(x.data_ptr()), None), "mm_index_create")
    rc = lib.mm_interpolate(idx, 1, 2, 3, 1, None, None, None, None, 1, None, 2 ** 31, None, 20, C.byref(prm), None,
                            None, None, None, None, None, 0, None)
    assert rc == -1 and b"2^31" in lib.mm_last_error()
    lib.mm_index_destroy(idx)


def test_ops_reject_hidden_copies(cuda):
    """Mis-strided / mis-aligned operands are errors, not silent .contiguous() / .clone() copies."""
    import torch
    from multimesh_b200 import ops

    nodes = torch.zeros((4, 27, 3), dtype=torch.float64, device=cuda)
    with pytest.raises(ValueError, match="not contiguous"):
        ops.element_geometry(nodes.transpose(0, 1))
    flat = torch.zeros((4 * 81 + 1,), dtype=torch.float64, device=cuda)
    with pytest.raises(ValueError, match="16-byte aligned"):
        ops.element_geometry(flat[1:].view(4, 27, 3))
