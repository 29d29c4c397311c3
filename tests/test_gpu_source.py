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


def test_host_pipeline_ramped_chunks_match_device_path(cuda, monkeypatch):
    """Default chunk schedule (small first chunks: 65 536, 131 072, ... points) on 300 k points == one device-side
    mm_interpolate over all of them, bit for bit, values and locations; and == the un-ramped schedule."""
    import torch
    from multimesh_b200 import ops

    monkeypatch.delenv("MM_HOST_CHUNK", raising=False)
    monkeypatch.delenv("MM_HOST_RAMP", raising=False)
    rng = np.random.default_rng(77)
    nodes, fields, pts = _case(2, 12, rng, 300_000, lo=-0.02, hi=1.02)
    src = ops.ResidentSource(nodes, fields, form="gll")
    vals, elem, xi, nf = src.interpolate_host(pts, 20, ops.V1(), want_location=True)
    monkeypatch.setenv("MM_HOST_RAMP", "0")
    vals0, elem0, xi0, nf0 = src.interpolate_host(pts, 20, ops.V1(), want_location=True)
    assert nf == nf0 and np.array_equal(vals, vals0) and np.array_equal(elem, elem0) and np.array_equal(xi, xi0)
    tn, tf, tp = (torch.from_numpy(a).to(cuda) for a in (nodes, fields, pts))
    cent, box = ops.element_geometry(tn)
    E, P, _ = nodes.shape
    index = ops.GridIndex(tn.view(E * P, 3)).prepare_sites()
    d_vals, d_elem, d_xi, _, d_nf = ops.interpolate(index, P, tn, cent, box, tf, tp, 20, ops.V1(),
                                                    presolve=ops.element_presolve(tn))
    assert nf == int(d_nf.item())
    assert np.array_equal(vals, d_vals.cpu().numpy()) and np.array_equal(elem, d_elem.cpu().numpy())
    assert np.array_equal(xi, d_xi.cpu().numpy())


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


@pytest.mark.parametrize("case", ["lattice_ties", "random_centroids", "clustered", "quads_2d", "outside", "big_coords"])
def test_tile_first_pass_matches_per_thread_kernel_and_oracle(cuda, oracle, case, monkeypatch):
    """The CTA-tile fp32-prefiltered first pass (knn_tile_kernel; writes certified PREFIXES of the canonical list,
    the pipeline re-runs what they do not resolve) must give the same final result as the per-thread exact kernels
    and as the oracle: identical target mesh (distances tied everywhere), sparse and crowded cells, 2-D, targets
    outside the source box, coordinates of O(6.4e6) with small elements; also with a tiny staging capacity, which
    forces the sub-tile split and the give-up (everything re-run) paths."""
    import torch
    from multimesh_b200 import ops

    rng = np.random.default_rng(abs(hash(case)) % 1000)
    dim, order, form = 3, 2, "gll"
    if case == "lattice_ties":
        nodes = meshgen.box_mesh((7, 6, 5), 2)
        pts = nodes.reshape(-1, 3).copy()  # every target coincides with a source GLL point: maximal ties
    elif case == "random_centroids":
        nodes = meshgen.box_mesh((9, 9, 9), 2, warp=0.04)
        pts = rng.uniform(0.0, 1.0, (20000, 3))
        form = "centroid"
    elif case == "clustered":
        # strongly graded element sizes: crowded and empty cells in one index
        nodes = meshgen.box_mesh((10, 10, 10), 2)
        nodes = nodes ** 3
        pts = rng.uniform(0.0, 1.0, (15000, 3)) ** 3
        form = "centroid"
    elif case == "quads_2d":
        dim = 2
        nodes = meshgen.box_mesh((30, 25), 2, warp=0.02)
        pts = np.concatenate([rng.uniform(-0.05, 1.05, (8000, 2)), nodes.reshape(-1, 2)[::3]])
    elif case == "outside":
        nodes = meshgen.box_mesh((6, 6, 6), 2, warp=0.02)
        pts = rng.uniform(-0.6, 1.6, (12000, 3))
    else:
        nodes = meshgen.box_mesh((8, 8, 8), 2, lo=[6.2e6, -4e4, -4e4], hi=[6.28e6, 4e4, 4e4], warp=0.01)
        pts = rng.uniform([6.2e6, -4e4, -4e4], [6.28e6, 4e4, 4e4], (15000, 3))
    E, P, _ = nodes.shape
    fields = rng.normal(size=(E, 3, P))
    if form == "gll":
        cands = (oracle.knn_bruteforce(nodes.reshape(-1, dim), pts, 20) // P).astype(np.int32)
    else:
        cands = oracle.knn_bruteforce(oracle.centroids(nodes), pts, 20)
    o_elem, o_xi, o_st, o_nf = oracle.locate(order, dim, nodes, pts, cands, oracle.V1())
    want = oracle.interp(order, dim, fields, o_elem, o_xi)
    tn, tf, tp = (torch.from_numpy(np.ascontiguousarray(a)).to(cuda) for a in (nodes, fields, pts))
    cent, box = ops.element_geometry(tn)
    pre = ops.element_presolve(tn)
    index = ops.GridIndex(tn.view(E * P, dim) if form == "gll" else cent)
    div = P if form == "gll" else 1
    for flag, cap in (("1", None), ("1", "200"), ("1", "40"), ("0", None)):
        monkeypatch.setenv("MM_KNN_TILE", flag)
        if cap:
            monkeypatch.setenv("MM_KT_CAP", cap)
        else:
            monkeypatch.delenv("MM_KT_CAP", raising=False)
        out, elem, xi, st, nf = ops.interpolate(index, div, tn, cent, box, tf, tp, 20, ops.V1(), presolve=pre)
        assert np.array_equal(elem.cpu().numpy(), o_elem), (case, flag)
        assert np.array_equal(st.cpu().numpy(), o_st), (case, flag)
        assert np.array_equal(xi.cpu().numpy(), o_xi) and int(nf.item()) == o_nf
        assert np.array_equal(out.cpu().numpy(), want)


@pytest.mark.parametrize("case", ["gll3d", "gll2d", "random_dups", "signed", "single"])
def test_unique_points_equals_numpy(cuda, case):
    """K4 (mm_unique_points): rows and inverse identical to np.unique(axis=0, return_inverse=True)."""
    import torch
    from multimesh_b200 import ops

    rng = np.random.default_rng(3)
    if case == "gll3d":
        pts = meshgen.box_mesh((9, 7, 8), 2, warp=0.03).reshape(-1, 3)
    elif case == "gll2d":
        pts = meshgen.box_mesh((21, 17), 4, warp=0.02).reshape(-1, 2)
    elif case == "random_dups":
        base = rng.uniform(-3.0, 5.0, (4000, 3))
        base[:, 0] = np.round(base[:, 0], 1)  # many equal x: the order is decided by y, then z
        base[::3, 1] = np.round(base[::3, 1], 0)
        pts = base[rng.integers(0, len(base), 30000)]
    elif case == "signed":
        pts = np.concatenate([rng.normal(size=(5000, 3)) * 1e6, np.zeros((5, 3)), -np.zeros((5, 3)),
                              np.array([[0.0, -0.0, 1.0], [-0.0, 0.0, 1.0], [1e-300, -1e-300, 0.0]])])
        pts = np.concatenate([pts, pts[:100]])
    else:
        pts = np.full((37, 3), 1.25)
    pts = np.ascontiguousarray(pts)
    u, inv = ops.unique_points(torch.from_numpy(pts).to(cuda))
    nu, ninv = np.unique(pts, return_inverse=True, axis=0)
    assert u.shape == nu.shape
    assert np.array_equal(u.cpu().numpy(), nu)  # (-0.0 == 0.0 compares equal, as in numpy's own grouping)
    assert np.array_equal(inv.cpu().numpy(), ninv.reshape(-1))
    assert np.array_equal(u.cpu().numpy()[inv.cpu().numpy()], pts)


def test_scatter_back_and_fluid_fixup_equal_numpy(cuda):
    import torch
    from multimesh_b200 import ops

    rng = np.random.default_rng(9)
    E, P, F, Nu = 50, 27, 5, 700
    vals = rng.normal(size=(Nu, F))
    inv = rng.integers(0, Nu, E * P).astype(np.int32)
    want = vals[inv].reshape(E, P, F).swapaxes(1, 2)
    got = ops.scatter_back(torch.from_numpy(vals).to(cuda), torch.from_numpy(inv).to(cuda), E, P)
    assert np.array_equal(got.cpu().numpy(), want)
    full = rng.normal(size=(E * P, F))
    got = ops.scatter_back(torch.from_numpy(full).to(cuda), None, E, P)
    assert np.array_equal(got.cpu().numpy(), full.reshape(E, P, F).swapaxes(1, 2))
    # fluid / solid repair (interpolator.py:829-841)
    new = want.copy()
    new[7, 4, 3] = 0.0   # a solid element that picked up a fluid value
    new[20, 4, :] = 0.0  # a fluid element (kept anyway)
    old = rng.normal(size=new.shape)
    fluid = np.zeros(E, dtype=bool)
    fluid[[3, 20]] = True
    ref = new.copy()
    ref[fluid] = old[fluid]
    for e in np.unique(np.where(ref[:, 4, :] == 0.0)[0]):
        if not fluid[e]:
            ref[e] = old[e]
    t = torch.from_numpy(new.copy()).to(cuda)
    ops.fluid_fixup_(t, torch.from_numpy(old).to(cuda), torch.from_numpy(fluid).to(cuda), 4)
    assert np.array_equal(t.cpu().numpy(), ref)


def test_gll_2_gll_cache_formats_round_trip(cuda, oracle, tmp_path):
    """stored_array: the first run writes the reference's elements.npy / coeffs.npy AND the compact (elem, xi)
    cache; re-runs from the compact cache (pure K3) and from the reference format (explicit-matrix gather) give the
    same file, bit for bit the first and within 1e-10 the second (different summation order)."""
    import multi_mesh.api as api
    from multimesh_b200.components import interpolator as itp
    from multimesh_b200.io.store import open_store, write_gll_model

    src = meshgen.box_mesh((6, 6, 6), 2, warp=0.02)
    tgt = meshgen.box_mesh((5, 5, 5), 2, lo=[0.02] * 3, hi=[0.98] * 3)
    a, store = str(tmp_path / "from.npz"), str(tmp_path / "stored")
    write_gll_model(a, src, meshgen.analytic_fields(src, NAMES), NAMES, np.zeros((216, 2)), ["fluid", "layer"])
    results = []
    for run in range(3):
        b = str(tmp_path / f"to{run}.npz")
        write_gll_model(b, tgt, np.zeros((125, 5, 27)), NAMES, np.zeros((125, 2)), ["fluid", "layer"])
        if run == 2:
            os.remove(os.path.join(store, itp.COMPACT_CACHE))  # force the reference format
        api.gll_2_gll(a, b, stored_array=store, gradient=True)
        with open_store(b, "r") as st:
            results.append(st.read("MODEL/data"))
        assert os.path.exists(os.path.join(store, "elements.npy")) and os.path.exists(os.path.join(store, "coeffs.npy"))
        if run < 2:
            assert os.path.exists(os.path.join(store, itp.COMPACT_CACHE))
    assert np.array_equal(results[0], results[1])
    assert np.max(np.abs(results[2] - results[0]) / np.abs(results[0])) <= 1e-10
