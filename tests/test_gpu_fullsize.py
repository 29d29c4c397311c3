"""
Parity at (near) benchmark size through size-independent properties (the oracle is too slow to
run on millions of points): identity interpolation, polynomial reproduction, linearity in the
fields, permutation invariance, agreement of the fused pipeline with the separate kernels, and
agreement of the two candidate forms where both must find the same owner.
Source: 64^3 order-2 hex mesh (262 144 elements), up to 7.1 M target points.
"""
import numpy as np
import pytest

from multimesh_b200 import meshgen

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def big(cuda):
    import torch
    from multimesh_b200 import ops

    nodes = meshgen.box_mesh((64, 64, 64), 2)
    E, P, _ = nodes.shape
    rng = np.random.default_rng(11)
    poly, coef = meshgen.polynomial_field(nodes, 2, rng)
    names = ["QKAPPA", "QMU", "RHO", "VP", "VS"]
    fields = meshgen.analytic_fields(nodes, names)
    fields[:, 0, :] = poly
    tn = torch.from_numpy(nodes).to(cuda)
    tf = torch.from_numpy(fields).to(cuda)
    cent, box = ops.element_geometry(tn)
    pre = ops.element_presolve(tn)
    return dict(nodes=nodes, tn=tn, tf=tf, fields=fields, cent=cent, box=box, pre=pre, coef=coef,
                gll=ops.GridIndex(tn.view(E * P, 3)), cen=ops.GridIndex(cent), P=P, E=E)


def test_identity_all_gll_points(cuda, big):
    """mesh -> same mesh: 7.1 M targets, every one on a node shared by up to 8 elements."""
    import torch
    from multimesh_b200 import ops

    pts = big["tn"].reshape(-1, 3)
    out, elem, xi, st, nf = ops.interpolate(big["gll"], big["P"], big["tn"], big["cent"], big["box"], big["tf"],
                                            pts, 20, ops.V1(), presolve=big["pre"])
    want = big["tf"].permute(0, 2, 1).reshape(-1, 5)
    rel = ((out - want).abs() / want.abs()).max().item()
    assert int(nf.item()) == 0 and rel <= 1e-10
    assert bool((st == 0).all().item())
    assert float(xi.abs().max().item()) <= 1.0 + 1e-12  # nodes sit at |xi| <= 1 of their owner


def test_polynomial_reproduction_linearity_permutation(cuda, big):
    import torch
    from multimesh_b200 import ops

    g = torch.Generator(device=cuda)
    g.manual_seed(5)
    N = 4_000_000
    pts = torch.rand((N, 3), dtype=torch.float64, device=cuda, generator=g)
    args = (big["tn"], big["cent"], big["box"])
    out, elem, xi, st, nf = ops.interpolate(big["gll"], big["P"], *args, big["tf"], pts, 20, ops.V1(),
                                            presolve=big["pre"])
    assert int(nf.item()) == 0
    # degree-2-per-axis polynomial is reproduced by order-2 elements
    coef = torch.from_numpy(big["coef"]).to(cuda)
    exact = torch.zeros(N, dtype=torch.float64, device=cuda)
    for i in range(3):
        for j in range(3):
            for k in range(3):
                exact += coef[i, j, k] * pts[:, 0] ** i * pts[:, 1] ** j * pts[:, 2] ** k
    assert float((out[:, 0] - exact).abs().max().item()) < 1e-11
    # linearity in the fields: interp(2 f + 3 g) = 2 interp(f) + 3 interp(g)
    comb = (2.0 * big["tf"][:, 3:4, :] + 3.0 * big["tf"][:, 2:3, :]).contiguous()
    oc = ops.interp_perm(comb, elem, xi, None)[:, 0]
    lin = 2.0 * out[:, 3] + 3.0 * out[:, 2]
    assert float(((oc - lin).abs() / lin.abs()).max().item()) < 1e-13
    # permuting the targets permutes the results (bit for bit)
    perm = torch.randperm(N, device=cuda, generator=g)
    out2, elem2, xi2, _, _ = ops.interpolate(big["gll"], big["P"], *args, big["tf"], pts[perm].contiguous(), 20,
                                             ops.V1(), presolve=big["pre"])
    assert torch.equal(out2, out[perm]) and torch.equal(elem2, elem[perm]) and torch.equal(xi2, xi[perm])
    # fused pipeline == separate kernels on a 1 M subset
    sub = pts[:1_000_000].contiguous()
    cands = big["gll"].query_idx(sub, 20, divisor=big["P"])
    e3, x3, s3, _ = ops.locate(*args, sub, cands, ops.V1(), presolve=big["pre"])
    assert torch.equal(e3, elem[:1_000_000]) and torch.equal(x3, xi[:1_000_000])
    assert torch.equal(ops.interp(big["tf"], e3, x3), out[:1_000_000])
    # interior points: the centroid form must find the same owner and the same xi
    _, e4, x4, _, _ = ops.interpolate(big["cen"], 1, *args, None, sub, 20, ops.V1(), presolve=big["pre"])
    inner = (x3.abs().max(dim=1).values < 0.95)
    assert torch.equal(e4[inner], e3[inner]) and torch.equal(x4[inner], x3[inner])


def test_layered_shell_54k_elements_order4_vs_oracle(cuda, oracle):
    """BASELINE config 3 at a size the oracle still finishes in seconds: cubed-sphere shell of 54 000 curved order-4
    elements (coordinates of O(6.4e6) m, mantle + lower crust + a 25 km upper crust), one index PER LAYER over the
    layer's centroids with layer-local element ids (components/interpolator.py:363-373), targets = all GLL points of
    a differently resolved shell with the same layer radii (257 k points); V1 (gll_2_gll_layered) and V2 with
    snap_to_nearest (gll_2_gll_layered_multi_two) -- elements, xi, status, failed counts and values bit-equal to the
    CPU oracle.  Exercises the occupied-volume cell size on thin shells, the CTA-tile K1, the grouped K2 with three
    Newton evaluations per point, and the element-centric K3 at order 4."""
    import torch
    from multimesh_b200 import ops

    radii = ((3480e3, 6291e3, 3), (6291e3, 6346e3, 2), (6346e3, 6371e3, 1))
    src_layers = [(r0, r1, nr, lid, 0) for (r0, r1, lid), nr in zip(radii, (8, 1, 1))]
    tgt_layers = [(r0, r1, nr, lid, 0) for (r0, r1, lid), nr in zip(radii, (5, 1, 1))]
    src, sel, _ = meshgen.shell_mesh(30, src_layers, 4)   # 6 * 30^2 * 10 = 54 000 elements
    tgt, tel, _ = meshgen.shell_mesh(7, tgt_layers, 4)    # 6 * 7^2 * 7 = 2 058 elements = 257 250 GLL points
    assert src.shape == (54000, 125, 3)
    names = ["QKAPPA", "QMU", "RHO", "VP", "VS"]
    fields = meshgen.analytic_fields(src / 6371000.0, names)
    total = 0
    for lid in (3, 2, 1):
        sm, tm = sel["layer"] == lid, tel["layer"] == lid
        nodes = np.ascontiguousarray(src[sm])
        f = np.ascontiguousarray(fields[sm])
        pts = np.ascontiguousarray(tgt[tm].reshape(-1, 3))
        tn, tf, tp = (torch.from_numpy(a).to(cuda) for a in (nodes, f, pts))
        cent, box = ops.element_geometry(tn)
        pre = ops.element_presolve(tn)
        index = ops.GridIndex(cent)
        o_cent = oracle.centroids(nodes)
        for spec, prm, k in ((ops.V1(), oracle.V1(), 20), (ops.V2(1.05, True), oracle.V2(1.05, True), 30)):
            out, elem, xi, st, nf = ops.interpolate(index, 1, tn, cent, box, tf, tp, k, spec, presolve=pre)
            cands = oracle.knn_ckdtree_canonical(o_cent, pts, k, pad=16)
            o_elem, o_xi, o_st, o_nf = oracle.locate(4, 3, nodes, pts, cands, prm)
            assert np.array_equal(elem.cpu().numpy(), o_elem), (lid, k)
            assert np.array_equal(st.cpu().numpy(), o_st) and int(nf.item()) == o_nf, (lid, k)
            assert np.array_equal(xi.cpu().numpy(), o_xi), (lid, k)
            assert np.array_equal(out.cpu().numpy(), oracle.interp(4, 3, f, o_elem, o_xi)), (lid, k)
        total += len(pts)
    assert total == 2058 * 125
