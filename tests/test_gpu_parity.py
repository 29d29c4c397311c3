"""
GPU parity tests (the parity tests proper): every CUDA kernel, called through the C-ABI via the
torch custom ops, against the CPU oracle on the same seeded inputs.

Bar (BASELINE.json north_star): element ownership and k-NN indices bit-exact; reference
coordinates within 1e-12 absolute; interpolated values within 1e-10 relative.  Because the
kernels and the oracle share one documented operation order, the tests additionally assert
bit-identity and report it -- the tolerance asserts are the contract, bit-identity the design goal.
"""
import numpy as np
import pytest

from multimesh_b200 import meshgen

pytestmark = pytest.mark.gpu

XI_ATOL = 1e-12
VAL_RTOL = 1e-10


def _t(a, dev, dtype=None):
    import torch

    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype).to(dev)


def _mesh(order, dim, n, warp, lo=None, hi=None):
    shape = (n,) * dim
    return meshgen.box_mesh(shape, order, lo=lo, hi=hi, warp=warp)


def _targets(rng, dim, n, lo=-0.03, hi=1.03):
    return rng.uniform(lo, hi, size=(n, dim))


# ---------------------------------------------------------------------------------------------
# K0
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("order,dim", [(1, 2), (2, 2), (4, 2), (1, 3), (2, 3), (4, 3)])
def test_element_geometry_bit_exact(cuda, oracle, order, dim):
    from multimesh_b200 import ops

    nodes = _mesh(order, dim, 5, 0.04)
    cent, box = ops.element_geometry(_t(nodes, cuda))
    assert np.array_equal(cent.cpu().numpy(), oracle.centroids(nodes))
    assert np.array_equal(box.cpu().numpy(), oracle.aabb(nodes))
    # and bit-equal to the reference's np.mean(points, axis=1) (salvus_mesh_reader.py:99-100)
    assert np.array_equal(cent.cpu().numpy(), np.mean(nodes, axis=1))


# ---------------------------------------------------------------------------------------------
# K1
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dim,k", [(3, 20), (3, 25), (3, 30), (2, 20), (3, 1), (3, 64)])
def test_knn_random_bit_exact(cuda, oracle, dim, k):
    from multimesh_b200 import ops

    rng = np.random.default_rng(10 + dim + k)
    data = rng.random((5000, dim))
    pts = _targets(rng, dim, 3000, -0.2, 1.2)
    ix = ops.GridIndex(_t(data, cuda))
    got = ix.query_idx(_t(pts, cuda), k).cpu().numpy()
    want = oracle.knn_bruteforce(data, pts, k)
    assert np.array_equal(got, want)
    dist, idx = ix.query(_t(pts, cuda), k)
    _, d2 = oracle.knn_bruteforce(data, pts, k, return_d2=True)
    assert np.array_equal(dist.cpu().numpy(), np.sqrt(d2))
    ix.close()


def test_knn_structured_ties_bit_exact(cuda, oracle):
    """Structured-grid centroids with targets ON mesh nodes: 8-way exact distance ties
    (SURVEY fact 5); the canonical order is (d2, index)."""
    from multimesh_b200 import ops

    nodes = _mesh(2, 3, 8, 0.0)
    cent = oracle.centroids(nodes)
    pts = nodes.reshape(-1, 3)[::7]
    ix = ops.GridIndex(_t(cent, cuda))
    got = ix.query_idx(_t(pts, cuda), 20).cpu().numpy()
    assert np.array_equal(got, oracle.knn_bruteforce(cent, pts, 20))


def test_knn_gll_point_form_duplicates(cuda, oracle):
    """gll_2_gll form: tree over ALL GLL points (shared nodes appear up to 8 times),
    idx // P -> element ids (interpolator.py:674-678,751-756)."""
    from multimesh_b200 import ops

    nodes = _mesh(2, 3, 6, 0.02)
    allp = nodes.reshape(-1, 3)
    rng = np.random.default_rng(3)
    pts = np.concatenate([_targets(rng, 3, 1500), allp[::11]])
    ix = ops.GridIndex(_t(allp, cuda))
    got = ix.query_idx(_t(pts, cuda), 20, divisor=27).cpu().numpy()
    want = oracle.knn_bruteforce(allp, pts, 20) // 27
    assert np.array_equal(got, want)


def test_knn_queries_far_outside_the_source(cuda, oracle):
    """Targets well outside the source's bounding box (a target mesh larger than the source): the ring
    search must stay exact and must terminate on the out-of-box distance, not after O(distance / cell)^3 rows."""
    import time

    import torch
    from multimesh_b200 import ops

    rng = np.random.default_rng(77)
    nodes = _mesh(2, 3, 12, 0.01)
    allp = nodes.reshape(-1, 3)
    far = np.concatenate([rng.uniform(-4.0, 5.0, (1500, 3)),             # up to 4 domain widths away, all octants
                          np.stack([rng.uniform(-60, 60, 300), rng.uniform(0, 1, 300), rng.uniform(0, 1, 300)], 1),
                          np.array([[1e6, 0.5, 0.5], [-1e6, -1e6, -1e6], [0.5, 0.5, 1e9]])])
    for data, div in ((oracle.centroids(nodes), 1), (allp, 27)):
        ix = ops.GridIndex(_t(data, cuda))
        for k in (4, 20):
            got = ix.query_idx(_t(far, cuda), k, divisor=div).cpu().numpy()
            assert np.array_equal(got, oracle.knn_bruteforce(data, far, k) // div)
    # the fused pipeline on the same targets (site pass + progressive search) and a timing guard
    tn = _t(nodes, cuda)
    cent, box = ops.element_geometry(tn)
    ix = ops.GridIndex(tn.view(-1, 3))
    big = _t(np.tile(far, (200, 1)), cuda)  # 360 k far targets
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out, elem, xi, st, nf = ops.interpolate(ix, 27, tn, cent, box, None, big, 20, ops.V1())
    torch.cuda.synchronize()
    assert time.perf_counter() - t0 < 20.0  # generous bound; the point is that it does not scale with the distance
    cands = oracle.knn_bruteforce(allp, far, 20) // 27
    e, x, _, _ = oracle.locate(2, 3, nodes, far, cands.astype(np.int32), oracle.V1())
    assert np.array_equal(elem[: len(far)].cpu().numpy(), e)


def test_non_finite_queries_fail_fast(cuda, oracle):
    """NaN / infinite target coordinates: no neighbours, the point is reported as failed (elem -1, zero row),
    and the other points of the batch are unaffected -- without a walk over the whole grid."""
    import time

    import torch
    from multimesh_b200 import ops

    nodes = _mesh(2, 3, 20, 0.01)
    fields = meshgen.analytic_fields(nodes, ["VP", "VS"])
    rng = np.random.default_rng(3)
    pts = rng.random((4000, 3))
    bad = np.arange(0, 4000, 97)
    pts[bad[0::3], 0] = np.nan
    pts[bad[1::3], 1] = np.inf
    pts[bad[2::3], 2] = -np.inf
    tn, tf = _t(nodes, cuda), _t(fields, cuda)
    cent, box = ops.element_geometry(tn)
    for data, div in ((tn.view(-1, 3), 27), (cent, 1)):
        ix = ops.GridIndex(data)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        idx = ix.query_idx(_t(pts, cuda), 20, divisor=div)
        out, elem, xi, st, nf = ops.interpolate(ix, div, tn, cent, box, tf, _t(pts, cuda), 20, ops.V2())
        torch.cuda.synchronize()
        assert time.perf_counter() - t0 < 10.0  # generous bound (first-use set-up included)
        idx, elem, out = idx.cpu().numpy(), elem.cpu().numpy(), out.cpu().numpy()
        assert (idx[bad] == -1).all() and (elem[bad] == -1).all() and (out[bad] == 0).all()
        assert int(nf.item()) == len(bad)
        good = np.setdiff1d(np.arange(4000), bad)
        want = oracle.knn_bruteforce(data.cpu().numpy(), pts[good], 20) // div
        assert np.array_equal(idx[good], want)
        e, x, _, _ = oracle.locate(2, 3, nodes, pts[good], want.astype(np.int32), oracle.V2())
        assert np.array_equal(elem[good], e)


def test_out_of_range_ids_are_treated_as_missing(cuda, oracle):
    """Candidate / element ids come from the caller through the C-ABI: ids outside [0, E) must behave like the
    -1 padding (skipped candidate, zero row), never read outside the arrays."""
    import torch
    from multimesh_b200 import ops

    nodes = _mesh(2, 3, 5, 0.02)
    E = nodes.shape[0]
    fields = meshgen.analytic_fields(nodes, ["VP", "VS", "RHO"])
    rng = np.random.default_rng(8)
    pts = rng.random((2000, 3))
    cands = oracle.knn_bruteforce(oracle.centroids(nodes), pts, 20).astype(np.int32)
    broken = cands.copy()
    broken[::3, 0] = E + 7          # first candidate replaced by garbage
    broken[1::3, 1] = 2**31 - 1
    clean = cands.copy()
    clean[::3, 0] = -1
    clean[1::3, 1] = -1
    tn, tf = _t(nodes, cuda), _t(fields, cuda)
    cent, box = ops.element_geometry(tn)
    for spec in (ops.V1(), ops.V2(snap_to_nearest=True)):
        e1, x1, s1, n1 = ops.locate(tn, cent, box, _t(pts, cuda), _t(broken, cuda), spec)
        e2, x2, s2, n2 = ops.locate(tn, cent, box, _t(pts, cuda), _t(clean, cuda), spec)
        assert torch.equal(e1, e2) and torch.equal(x1, x2) and torch.equal(s1, s2)
    elem = e2.clone()
    elem[::5] = E            # out of range -> zero row
    elem[1::5] = -3
    out = ops.interp(tf, elem, x2).cpu().numpy()
    ref = elem.clone()
    ref[::5] = -1
    ref[1::5] = -1
    assert np.array_equal(out, ops.interp(tf, ref, x2).cpu().numpy())
    assert (out[::5] == 0).all() and (out[1::5] == 0).all() and (out[2::5] != 0).all()
    perm = torch.randperm(len(pts), device=cuda).to(torch.int32)
    assert torch.equal(ops.interp_perm(tf, elem, x2, perm)[perm.long()], torch.from_numpy(out).to(cuda))


def test_knn_edge_cases(cuda, oracle):
    from multimesh_b200 import ops
    import torch

    rng = np.random.default_rng(5)
    data = rng.random((7, 3))
    pts = rng.random((33, 3))
    ix = ops.GridIndex(_t(data, cuda))
    got = ix.query_idx(_t(pts, cuda), 20).cpu().numpy()  # k > M: padded with -1
    assert np.array_equal(got, oracle.knn_bruteforce(data, pts, 20))
    assert (got[:, 7:] == -1).all()
    # empty query
    assert ix.query_idx(torch.empty((0, 3), dtype=torch.float64, device=cuda), 5).shape == (0, 5)
    # all data points identical / coplanar data
    same = np.tile(rng.random((1, 3)), (50, 1))
    got = ops.GridIndex(_t(same, cuda)).query_idx(_t(pts, cuda), 10).cpu().numpy()
    assert np.array_equal(got, oracle.knn_bruteforce(same, pts, 10))
    flat = rng.random((400, 3))
    flat[:, 2] = 0.25
    got = ops.GridIndex(_t(flat, cuda)).query_idx(_t(pts, cuda), 10).cpu().numpy()
    assert np.array_equal(got, oracle.knn_bruteforce(flat, pts, 10))
    # strongly non-uniform density
    clus = np.concatenate([rng.random((2000, 3)) * 0.01, rng.random((200, 3))])
    q = np.concatenate([rng.random((200, 3)) * 0.01, rng.random((200, 3))])
    got = ops.GridIndex(_t(clus, cuda)).query_idx(_t(q, cuda), 20).cpu().numpy()
    assert np.array_equal(got, oracle.knn_bruteforce(clus, q, 20))


# ---------------------------------------------------------------------------------------------
# K2
# ---------------------------------------------------------------------------------------------
def _variants(oracle):
    from multimesh_b200 import ops

    return [
        ("V1", ops.V1(), oracle.V1()),
        ("V2", ops.V2(), oracle.V2()),
        ("V2snap", ops.V2(1.05, True), oracle.V2(1.05, True)),
        ("V3", ops.V3(), oracle.V3()),
        ("V4", ops.V4(), oracle.V4()),
        ("V5", ops.V5(), oracle.V5()),
    ]


def _locate_both(cuda, oracle, nodes, pts, cands, spec, prm, use_pre=True):
    """use_pre: Newton starts from the affine pre-solve (K0) on both sides; else from xi = 0."""
    from multimesh_b200 import ops

    order = round(nodes.shape[1] ** (1.0 / nodes.shape[2])) - 1
    dim = nodes.shape[2]
    tn = _t(nodes, cuda)
    cent, box = ops.element_geometry(tn)
    pre = ops.element_presolve(tn) if use_pre else None
    if use_pre:
        assert np.array_equal(pre.cpu().numpy(), oracle.presolve(nodes)), "presolve not bit-identical"
    elem, xi, status, nfail = ops.locate(tn, cent, box, _t(pts, cuda), _t(cands, cuda), spec, presolve=pre)
    o_elem, o_xi, o_status, o_nfail = oracle.locate(order, dim, nodes, pts, cands, prm, pre=use_pre)
    return (elem.cpu().numpy(), xi.cpu().numpy(), status.cpu().numpy(), int(nfail.item()),
            o_elem, o_xi, o_status, o_nfail)


@pytest.mark.parametrize("order,dim,n,warp", [
    (1, 3, 6, 0.03), (2, 3, 6, 0.03), (4, 3, 4, 0.03), (2, 3, 7, 0.0),
    (1, 2, 9, 0.03), (2, 2, 9, 0.03), (4, 2, 6, 0.03),
])
def test_locate_all_variants(cuda, oracle, order, dim, n, warp):
    rng = np.random.default_rng(100 * order + dim)
    nodes = _mesh(order, dim, n, warp)
    # inside points, points outside the mesh (fallbacks), and points exactly on nodes (ties)
    pts = np.concatenate([_targets(rng, dim, 1501, -0.08, 1.08), nodes.reshape(-1, dim)[::5]])
    cands = oracle.knn_bruteforce(oracle.centroids(nodes), pts, 20)
    ref_xi = {}
    for use_pre in (True, False):
        for name, spec, prm in _variants(oracle):
            elem, xi, st, nf, o_elem, o_xi, o_st, o_nf = _locate_both(cuda, oracle, nodes, pts, cands, spec, prm,
                                                                      use_pre)
            assert np.array_equal(elem, o_elem), name          # ownership: bit-exact
            assert np.array_equal(st, o_st), name
            assert nf == o_nf == int((o_elem < 0).sum()), name
            assert np.max(np.abs(xi - o_xi), initial=0.0) <= XI_ATOL, name
            assert np.array_equal(xi, o_xi), f"{name}: xi not bit-identical"
            if use_pre:
                ref_xi[name] = (elem, xi, st)
            else:
                # the two Newton starts agree on ownership and to roundoff on xi for every point
                # accepted in the candidate loop (fallbacks that rank candidates by |xi| may
                # legitimately flip on roundoff-level ties)
                acc = (st == 0) & (ref_xi[name][2] == 0)
                assert acc.sum() > 0.5 * len(st), name
                assert np.array_equal(ref_xi[name][0][acc], elem[acc]), name
                assert np.max(np.abs(ref_xi[name][1][acc] - xi[acc]), initial=0.0) <= XI_ATOL, name
    # sanity: most points are accepted by the plain V2 search
    assert (o_elem >= 0).mean() > 0.5


def test_locate_gll_point_form_duplicates_and_padding(cuda, oracle):
    rng = np.random.default_rng(77)
    nodes = _mesh(2, 3, 5, 0.02)
    pts = _targets(rng, 3, 700)
    cands = (oracle.knn_bruteforce(nodes.reshape(-1, 3), pts, 20) // 27).astype(np.int32)
    cands[::9, 3:] = -1  # ragged candidate lists
    cands[5] = -1        # a point with no candidates at all
    from multimesh_b200 import ops

    for spec, prm in [(ops.V1(), oracle.V1()), (ops.V3(), oracle.V3())]:
        elem, xi, st, nf, o_elem, o_xi, o_st, o_nf = _locate_both(cuda, oracle, nodes, pts, cands, spec, prm)
        assert np.array_equal(elem, o_elem) and np.array_equal(st, o_st) and nf == o_nf
        assert np.array_equal(xi, o_xi)
    assert o_elem[5] == -1


def test_locate_last_element_alignment(cuda, oracle):
    """The last element of an odd-sized mesh ends the array: exercises the plain-load tail path
    next to the bulk-copy path (3-D order 2/4 element blocks are 8 mod 16 bytes)."""
    from multimesh_b200 import ops

    for order in (2, 4):
        nodes = _mesh(order, 3, 3, 0.01)  # 27 elements: last id 26 is even
        assert nodes.shape[0] % 2 == 1
        rng = np.random.default_rng(order)
        pts = _targets(rng, 3, 257, 0.6, 1.0)
        cands = np.tile(np.array([26, 25, 13], dtype=np.int32), (len(pts), 1))
        elem, xi, st, nf, o_elem, o_xi, o_st, o_nf = _locate_both(
            cuda, oracle, nodes, pts, cands, ops.V2(), oracle.V2())
        assert np.array_equal(elem, o_elem) and np.array_equal(xi, o_xi)
        assert (o_elem == 26).any()


# ---------------------------------------------------------------------------------------------
# K3
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("order,dim,F", [
    (1, 3, 5), (2, 3, 5), (4, 3, 5), (2, 3, 8), (4, 3, 1), (2, 3, 3), (1, 3, 11),
    (1, 2, 3), (2, 2, 3), (4, 2, 3), (4, 2, 8),
])
def test_interp_bit_exact(cuda, oracle, order, dim, F):
    from multimesh_b200 import ops

    rng = np.random.default_rng(1000 + 10 * order + F)
    n = 3
    nodes = _mesh(order, dim, n, 0.0)
    E, P, _ = nodes.shape
    fields = rng.normal(size=(E, F, P)) * 1000.0
    N = 2049  # ragged: not a multiple of 32
    elem = rng.integers(0, E, size=N).astype(np.int32)
    elem[::13] = E - 1  # hit the array tail
    elem[5::17] = -1    # failed points -> zero rows
    xi = rng.uniform(-1.04, 1.04, size=(N, dim))
    out = ops.interp(_t(fields, cuda), _t(elem, cuda), _t(xi, cuda)).cpu().numpy()
    want = oracle.interp(order, dim, fields, elem, xi)
    scale = np.abs(fields).max() * P
    assert np.max(np.abs(out - want)) <= VAL_RTOL * scale
    assert np.array_equal(out, want), "K3 not bit-identical to the oracle"
    assert (out[elem < 0] == 0).all()
    c = ops.coeffs(_t(elem, cuda), _t(xi, cuda), order).cpu().numpy()
    assert np.array_equal(c, oracle.coeffs(order, dim, elem, xi))
    g = ops.gather_coeffs(_t(fields, cuda), _t(elem, cuda), _t(c, cuda)).cpu().numpy()
    ref = np.sum(fields[np.maximum(elem, 0)] * c[:, None, :], axis=2)  # numpy, as interpolator.py:822
    assert np.max(np.abs(g - ref)) <= VAL_RTOL * scale
    assert np.max(np.abs(out - ref)) <= VAL_RTOL * scale


def test_interp_empty(cuda):
    import torch
    from multimesh_b200 import ops

    f = torch.zeros((4, 5, 27), dtype=torch.float64, device=cuda)
    out = ops.interp(f, torch.empty((0,), dtype=torch.int32, device=cuda),
                     torch.empty((0, 3), dtype=torch.float64, device=cuda))
    assert out.shape == (0, 5)


# ---------------------------------------------------------------------------------------------
# whole pipeline: k-NN -> locate -> gather, polynomial reproduction and identity interpolation
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("order", [1, 2, 4])
def test_pipeline_polynomial_reproduction(cuda, oracle, order):
    from multimesh_b200 import ops

    rng = np.random.default_rng(order)
    nodes = _mesh(order, 3, 5, 0.0)
    v, coef = meshgen.polynomial_field(nodes, order, rng)
    fields = np.ascontiguousarray(v[:, None, :])
    pts = rng.uniform(0.0, 1.0, size=(4000, 3))
    tn = _t(nodes, cuda)
    cent, box = ops.element_geometry(tn)
    cands = ops.GridIndex(cent).query_idx(_t(pts, cuda), 20)
    elem, xi, st, nf = ops.locate(tn, cent, box, _t(pts, cuda), cands, ops.V1(),
                                  presolve=ops.element_presolve(tn))
    out = ops.interp(_t(fields, cuda), elem, xi).cpu().numpy()[:, 0]
    exact, _ = meshgen.polynomial_field(pts[None, :, :], order, rng) if False else (None, None)
    exact = np.zeros(len(pts))
    for idx in np.ndindex(*coef.shape):
        exact += coef[idx] * pts[:, 0] ** idx[0] * pts[:, 1] ** idx[1] * pts[:, 2] ** idx[2]
    assert int(nf.item()) == 0
    assert np.max(np.abs(out - exact)) < 1e-11


def test_pipeline_identity_mesh(cuda, oracle):
    """mesh -> same mesh returns the fields (every target sits on a node shared by up to 8
    elements: exercises the tie-break)."""
    from multimesh_b200 import ops

    nodes = _mesh(2, 3, 6, 0.02)
    names = ["QKAPPA", "QMU", "RHO", "VP", "VS"]
    fields = meshgen.analytic_fields(nodes, names)
    pts = nodes.reshape(-1, 3)
    tn = _t(nodes, cuda)
    cent, box = ops.element_geometry(tn)
    cands = ops.GridIndex(cent).query_idx(_t(pts, cuda), 20)
    elem, xi, st, nf = ops.locate(tn, cent, box, _t(pts, cuda), cands, ops.V1(),
                                  presolve=ops.element_presolve(tn))
    out = ops.interp(_t(fields, cuda), elem, xi).cpu().numpy()
    want = np.swapaxes(fields, 1, 2).reshape(-1, len(names))
    assert np.max(np.abs(out - want) / np.abs(want)) < VAL_RTOL
    o_cands = oracle.knn_bruteforce(oracle.centroids(nodes), pts, 20)
    assert np.array_equal(cands.cpu().numpy(), o_cands)
    o_elem, o_xi, _, _ = oracle.locate(2, 3, nodes, pts, o_cands, oracle.V1())
    assert np.array_equal(elem.cpu().numpy(), o_elem)


# ---------------------------------------------------------------------------------------------
# order-1 nodal (exodus) path
# ---------------------------------------------------------------------------------------------
def test_trilinear_bit_exact(cuda, oracle):
    from multimesh_b200 import ops

    rng = np.random.default_rng(8)
    points, conn = meshgen.hex8_mesh((7, 6, 5), warp=0.03)
    perm = np.argsort([0, 3, 2, 1, 4, 5, 6, 7])
    connC = np.ascontiguousarray(conn[:, perm])
    cent = ops.centroid_conn(_t(conn, cuda), _t(points, cuda))
    assert np.array_equal(cent.cpu().numpy(), oracle.centroid_conn(conn, points))
    q = np.concatenate([rng.uniform(-0.05, 1.05, (3000, 3)), points[::3]])
    nn = oracle.knn_bruteforce(cent.cpu().numpy(), q, 20).astype(np.int64)
    nfail, enc, w = ops.trilinear(_t(nn, cuda), _t(connC, cuda), _t(points, cuda), _t(q, cuda))
    o_nf, o_enc, o_w = oracle.trilinear_interpolator(20, nn, connC, points, q)
    assert int(nfail.item()) == o_nf
    assert np.array_equal(enc.cpu().numpy(), o_enc)
    assert np.array_equal(w.cpu().numpy(), o_w)
    if oracle.ref_lib() is not None:  # the compiled reference itself
        r_nf, r_enc, r_w = oracle.ref_trilinear_interpolator(20, nn, connC, points, q)
        assert r_nf == int(nfail.item())
        assert np.array_equal(enc.cpu().numpy(), r_enc) and np.array_equal(w.cpu().numpy(), r_w)
    param = rng.normal(size=(7, len(points)))  # 4 + 2 + 1: every chunk width of the field loop
    vals = ops.gather_nodal(_t(param, cuda), enc, w).cpu().numpy()
    want = np.sum(param[:, o_enc] * o_w, axis=2)
    assert np.max(np.abs(vals - want)) < 1e-12


@pytest.mark.parametrize("shape,warp,nq,k,env", [
    ((7, 6, 5), 0.03, 3000, 20, {}),                        # few cells: per-thread first pass
    ((24, 24, 24), 0.03, 60000, 20, {}),                    # CTA-tile first pass
    ((24, 24, 24), 0.12, 60000, 20, {"MM_RERUN_CHUNK": "700"}),  # strong warp: many re-runs, several rounds
    ((24, 24, 24), 0.03, 60000, 20, {"MM_KNN_TILE": "0"}),
    ((9, 9, 9), 0.03, 5000, 4, {}),                         # k <= 4: no prefix, complete semantics at once
    ((2, 2, 2), 0.0, 4000, 20, {}),                         # k > number of elements: -1 padding
])
def test_trilinear_indexed_matches_knn_plus_trilinear(cuda, oracle, monkeypatch, shape, warp, nq, k, env):
    """mm_trilinear_indexed (sort, 4-prefix, prefix search, re-run) == mm_knn(k) + mm_trilinear == oracle == compiled
    reference C: enclosing node ids, weights and the failed count, bit for bit -- including points outside the mesh
    (failed: rows stay zero) and points that need the second-chance candidate."""
    import torch
    from multimesh_b200 import ops

    for name, val in env.items():
        monkeypatch.setenv(name, val)
    rng = np.random.default_rng(81)
    points, conn = meshgen.hex8_mesh(shape, warp=warp)
    connC = np.ascontiguousarray(conn[:, np.argsort([0, 3, 2, 1, 4, 5, 6, 7])])
    t_conn, t_connC, t_points = _t(conn, cuda), _t(connC, cuda), _t(points, cuda)
    cent = ops.centroid_conn(t_conn, t_points)
    out = 0.08 if min(shape) >= 5 else 0.3  # coarse mesh: the second chance reaches far (|xi| < 1.5)
    q = np.concatenate([rng.uniform(-out, 1 + out, (nq, 3)), points[::3], cent.cpu().numpy()[::5]])
    t_q = _t(q, cuda)
    index = ops.GridIndex(cent)
    nf1, enc1, w1 = ops.trilinear_indexed(index, t_connC, t_points, t_q, k)
    nn = index.query_idx(t_q, k).to(torch.int64)
    nf0, enc0, w0 = ops.trilinear(nn, t_connC, t_points, t_q)
    assert int(nf1.item()) == int(nf0.item())
    assert torch.equal(enc1, enc0) and torch.equal(w1, w0)
    assert 0 < int(nf1.item()) < len(q)  # the box around the mesh: some points fail, most do not
    sel = rng.choice(len(q), min(len(q), 6000), replace=False)
    nn_h = nn.cpu().numpy()[sel]
    o_nf, o_enc, o_w = oracle.trilinear_interpolator(k, nn_h, connC, points, q[sel])
    assert np.array_equal(enc1.cpu().numpy()[sel], o_enc) and np.array_equal(w1.cpu().numpy()[sel], o_w)
    if oracle.ref_lib() is not None and (nn_h >= 0).all():
        r_nf, r_enc, r_w = oracle.ref_trilinear_interpolator(k, nn_h, connC, points, q[sel])
        assert r_nf == o_nf
        assert np.array_equal(enc1.cpu().numpy()[sel], r_enc) and np.array_equal(w1.cpu().numpy()[sel], r_w)


def test_legacy_host_symbols(cuda, oracle):
    """helpers.load_lib()-style ctypes calls on HOST numpy buffers (helpers.py:43-81)."""
    import ctypes as C
    from multimesh_b200 import _lib

    lib = _lib.load_lib()
    rng = np.random.default_rng(9)
    points, conn = meshgen.hex8_mesh((4, 4, 4), warp=0.02)
    cent = np.zeros((len(conn), 3))
    lib.centroid(3, len(conn), 8, conn.ctypes.data_as(C.c_void_p), points.ctypes.data_as(C.c_void_p),
                 cent.ctypes.data_as(C.c_void_p))
    assert np.array_equal(cent, oracle.centroid_conn(conn, points))
    connC = np.ascontiguousarray(conn[:, np.argsort([0, 3, 2, 1, 4, 5, 6, 7])])
    q = rng.uniform(0.0, 1.0, (500, 3))
    nn = oracle.knn_bruteforce(cent, q, 20).astype(np.int64)
    enc = np.zeros((len(q), 8), dtype=np.int64)
    w = np.zeros((len(q), 8))
    nf = lib.triLinearInterpolator(20, len(q), nn.ctypes.data_as(C.c_void_p),
                                   connC.ctypes.data_as(C.c_void_p), enc.ctypes.data_as(C.c_void_p),
                                   points.ctypes.data_as(C.c_void_p), w.ctypes.data_as(C.c_void_p),
                                   q.ctypes.data_as(C.c_void_p))
    o_nf, o_enc, o_w = oracle.trilinear_interpolator(20, nn, connC, points, q)
    assert nf == o_nf == 0
    assert np.array_equal(enc, o_enc) and np.array_equal(w, o_w)


def test_interpolate_host_entry_point(cuda, oracle):
    import ctypes as C
    from multimesh_b200 import _lib, ops

    lib = _lib.load_lib()
    rng = np.random.default_rng(12)
    nodes = _mesh(2, 3, 5, 0.02)
    fields = meshgen.analytic_fields(nodes, ["QKAPPA", "QMU", "RHO", "VP", "VS"])
    pts = _targets(rng, 3, 999, 0.0, 1.0)
    for form in (0, 1):
        vals = np.zeros((len(pts), 5))
        elem = np.zeros(len(pts), dtype=np.int32)
        xi = np.zeros((len(pts), 3))
        nf = C.c_int64(-1)
        prm = ops.V1().to_c()
        rc = lib.mm_interpolate_host(2, 3, nodes.shape[0], nodes.ctypes.data_as(C.c_void_p), 5,
                                     fields.ctypes.data_as(C.c_void_p), len(pts),
                                     pts.ctypes.data_as(C.c_void_p), 20, form, C.byref(prm),
                                     vals.ctypes.data_as(C.c_void_p), elem.ctypes.data_as(C.c_void_p),
                                     xi.ctypes.data_as(C.c_void_p), C.byref(nf))
        assert rc == 0, lib.mm_last_error()
        if form == 0:
            cands = oracle.knn_bruteforce(oracle.centroids(nodes), pts, 20)
        else:
            cands = oracle.knn_bruteforce(nodes.reshape(-1, 3), pts, 20) // 27
        o_elem, o_xi, _, o_nf = oracle.locate(2, 3, nodes, pts, cands, oracle.V1())
        assert np.array_equal(elem, o_elem) and np.array_equal(xi, o_xi) and nf.value == o_nf
        assert np.array_equal(vals, oracle.interp(2, 3, fields, o_elem, o_xi))


def test_cpu_tensors_are_rejected(cuda):
    import torch
    from multimesh_b200 import ops

    with pytest.raises(Exception):
        ops.interp(torch.zeros((1, 1, 8), dtype=torch.float64), torch.zeros(1, dtype=torch.int32),
                   torch.zeros((1, 3), dtype=torch.float64))


# ---------------------------------------------------------------------------------------------
# small-k (register list) k-NN and the fused, progressive pipeline
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dim,k", [(3, 2), (3, 4), (3, 5), (3, 8), (2, 3), (2, 8)])
def test_knn_small_k_register_list(cuda, oracle, dim, k):
    from multimesh_b200 import ops

    rng = np.random.default_rng(50 + dim + k)
    nodes = _mesh(2, dim, 6, 0.0)
    data = np.concatenate([nodes.reshape(-1, dim), rng.random((500, dim))])  # duplicates + random
    pts = np.concatenate([_targets(rng, dim, 1500, -0.1, 1.1), nodes.reshape(-1, dim)[::13]])
    got = ops.GridIndex(_t(data, cuda)).query_idx(_t(pts, cuda), k).cpu().numpy()
    assert np.array_equal(got, oracle.knn_bruteforce(data, pts, k))


@pytest.mark.parametrize("order,dim,form", [(2, 3, "gll"), (2, 3, "centroid"), (4, 3, "centroid"), (2, 2, "gll"),
                                            (1, 3, "centroid")])
def test_fused_pipeline_equals_separate_kernels(cuda, oracle, order, dim, form):
    """mm_interpolate (spatial sort + progressive search) == mm_knn -> mm_locate -> mm_interp,
    for every location variant, including points that need the full-k re-run and fallbacks."""
    from multimesh_b200 import ops

    rng = np.random.default_rng(7 * order + dim)
    nodes = _mesh(order, dim, 5 if order == 4 else 7, 0.03)
    E, P, _ = nodes.shape
    F = 5
    fields = rng.normal(size=(E, F, P)) * 100.0
    pts = np.concatenate([_targets(rng, dim, 3001, -0.1, 1.1), nodes.reshape(-1, dim)[::7]])
    tn, tf, tp = _t(nodes, cuda), _t(fields, cuda), _t(pts, cuda)
    cent, box = ops.element_geometry(tn)
    pre = ops.element_presolve(tn)
    if form == "gll":
        index, div = ops.GridIndex(tn.view(E * P, dim)), P
        data = nodes.reshape(-1, dim)
    else:
        index, div = ops.GridIndex(cent), 1
        data = oracle.centroids(nodes)
    for k in (20, 6):
        cands = (oracle.knn_bruteforce(data, pts, k) // div).astype(np.int32)
        for name, spec, prm in _variants(oracle):
            use_pre = name not in ("V3", "V5")  # exercise both Newton starts
            out, elem, xi, st, nf = ops.interpolate(index, div, tn, cent, box, tf, tp, k, spec,
                                                    presolve=pre if use_pre else None)
            o_elem, o_xi, o_st, o_nf = oracle.locate(order, dim, nodes, pts, cands, prm, pre=use_pre)
            assert np.array_equal(elem.cpu().numpy(), o_elem), (name, k)
            assert np.array_equal(st.cpu().numpy(), o_st), (name, k)
            assert np.array_equal(xi.cpu().numpy(), o_xi), (name, k)
            assert int(nf.item()) == o_nf, (name, k)
            assert np.array_equal(out.cpu().numpy(), oracle.interp(order, dim, fields, o_elem, o_xi)), (name, k)
    # locate-only mode, and no location outputs
    out, elem, xi, st, nf = ops.interpolate(index, div, tn, cent, box, None, tp, 20, ops.V1())
    assert out.numel() == 0 and elem.shape[0] == len(pts)
    out2, e2, _, _, _ = ops.interpolate(index, div, tn, cent, box, tf, tp, 20, ops.V1(), want_location=False)
    assert e2.numel() == 0 and out2.shape == (len(pts), F)


@pytest.mark.parametrize("order,dim,F", [(2, 3, 5), (4, 3, 5), (4, 3, 8), (1, 3, 5), (2, 2, 3), (2, 3, 40)])
def test_interp_coherent_variant_bit_exact(cuda, oracle, order, dim, F):
    """mm_interp_perm: warp-level de-duplication of element blocks; incoherent input (up to 32
    distinct elements per warp -> several rounds), coherent input, failed points, permutation."""
    import torch
    from multimesh_b200 import ops

    rng = np.random.default_rng(2000 + 10 * order + F)
    nodes = _mesh(order, dim, 4, 0.0)
    E, P, _ = nodes.shape
    fields = rng.normal(size=(E, F, P)) * 1000.0
    N = 3001
    for coherent in (False, True):
        elem = rng.integers(0, E, size=N).astype(np.int32)
        if coherent:
            elem = np.sort(elem)
        elem[::13] = E - 1
        elem[5::17] = -1
        xi = rng.uniform(-1.04, 1.04, size=(N, dim))
        perm = rng.permutation(N).astype(np.int32)
        want = oracle.interp(order, dim, fields, elem, xi)
        got = ops.interp_perm(_t(fields, cuda), _t(elem, cuda), _t(xi, cuda), None).cpu().numpy()
        assert np.array_equal(got, want)
        got = ops.interp_perm(_t(fields, cuda), _t(elem, cuda), _t(xi, cuda), _t(perm, cuda)).cpu().numpy()
        assert np.array_equal(got[perm], want)


@pytest.mark.parametrize("order,dim,F,nelem,npts", [
    (4, 3, 5, 3, 4000),    # one group of 25 lanes; hundreds of points per element: many chunks
    (4, 3, 8, 3, 3000),    # TTI: two field passes (6 + 2) AND several chunks -> Lagrange values recomputed per pass
    (2, 3, 5, 4, 5000),    # two groups of 15 lanes
    (2, 3, 13, 3, 2000),   # F > 10: two passes at order 2
    (1, 3, 5, 5, 3000),    # three groups of 10 lanes
    (1, 3, 40, 3, 500),    # three passes
    (2, 2, 3, 9, 4000),    # quads: one group of 9 lanes x 3 groups
    (4, 2, 7, 6, 3000),    # quads, two passes (6 + 1)
    (2, 3, 2, 4, 3000),    # few fields: five groups of 6 lanes (five elements per warp)
    (1, 3, 1, 5, 3000),    # sixteen groups of 2 lanes
    (4, 3, 1, 3, 2000),    # six groups of 5 lanes
    (1, 2, 1, 8, 2000),    # quads, sixteen groups
])
def test_pipeline_element_centric_gather(cuda, oracle, order, dim, F, nelem, npts, monkeypatch):
    """K3 of mm_interpolate in its element-centric form (mm_interp_elem.cu: points grouped by element, field slabs in
    registers, canonical operation order redistributed over lanes): bit-identical to the oracle's gather and to the
    point-order tile kernel; elements with zero, one and hundreds of points, points outside the mesh (zero rows)."""
    from multimesh_b200 import ops

    rng = np.random.default_rng(100 * order + 10 * dim + F)
    nodes = _mesh(order, dim, nelem, 0.02)
    E, P, _ = nodes.shape
    fields = rng.normal(size=(E, F, P)) * 1000.0
    # clustered targets: most in a corner (some elements get hundreds of points, others none), some outside the mesh
    pts = np.concatenate([rng.uniform(0.0, 0.45, (npts, dim)), rng.uniform(-0.3, 1.3, (npts // 10, dim)),
                          nodes.reshape(-1, dim)[::5]])
    tn, tf, tp = _t(nodes, cuda), _t(fields, cuda), _t(pts, cuda)
    cent, box = ops.element_geometry(tn)
    pre = ops.element_presolve(tn)
    index = ops.GridIndex(cent)
    cands = oracle.knn_bruteforce(oracle.centroids(nodes), pts, 20)
    seen_failed = seen_crowded = False
    for spec, prm in ((ops.V3(), oracle.V3()), (ops.V1(), oracle.V1())):
        o_elem, o_xi, o_st, o_nf = oracle.locate(order, dim, nodes, pts, cands, prm)
        seen_failed |= bool((o_elem < 0).any())
        seen_crowded |= bool(np.bincount(o_elem[o_elem >= 0], minlength=E).max() > 64)
        want = oracle.interp(order, dim, fields, o_elem, o_xi)
        res = {}
        for mode in (None, "t"):
            if mode:
                monkeypatch.setenv("MM_INTERP_MODE", mode)
            else:
                monkeypatch.delenv("MM_INTERP_MODE", raising=False)
            out, elem, xi, st, nf = ops.interpolate(index, 1, tn, cent, box, tf, tp, 20, spec, presolve=pre)
            assert np.array_equal(elem.cpu().numpy(), o_elem) and np.array_equal(st.cpu().numpy(), o_st)
            assert np.array_equal(xi.cpu().numpy(), o_xi) and int(nf.item()) == o_nf
            assert np.array_equal(out.cpu().numpy(), want), mode
            res[mode] = out
        out2, e2, _, _, _ = ops.interpolate(index, 1, tn, cent, box, tf, tp, 20, spec, presolve=pre, want_location=False)
        assert e2.numel() == 0 and np.array_equal(out2.cpu().numpy(), want)
    assert seen_failed and seen_crowded  # the cases this test is about did occur
