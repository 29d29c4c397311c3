"""
GPU tests of the reference-facing drivers (multi_mesh.api / components.interpolator mirrors):
each driver is run on synthetic meshes in the reference's file layout and compared with the
same workflow assembled from the CPU oracle, following the reference's own steps
(unique points -> k-NN -> locate -> weights -> gather -> scatter back / reshape).
"""
import os

import numpy as np
import pytest

from oracle import np_oracle

from multimesh_b200 import meshgen, utils
from multimesh_b200.components.salvus_mesh_reader import SalvusMesh
from multimesh_b200.io.exodus import Exodus
from multimesh_b200.io.store import open_store, write_gll_model

pytestmark = pytest.mark.gpu
ISO = ["QKAPPA", "QMU", "RHO", "VP", "VS"]


def _write(path, nodes, names, fluid=None, layer=None):
    data = meshgen.analytic_fields(nodes, names)
    E = nodes.shape[0]
    ed = np.stack([np.zeros(E) if fluid is None else fluid, np.ones(E) if layer is None else layer], axis=1)
    write_gll_model(path, nodes, data, names, ed, ["fluid", "layer"])
    return data


def _oracle_gll_2_gll(oracle, src_nodes, src_data, tgt_nodes, k):
    order = round(src_nodes.shape[1] ** (1.0 / src_nodes.shape[2])) - 1
    dim = src_nodes.shape[2]
    P = src_nodes.shape[1]
    uniq, recon = np_oracle.unique_points(tgt_nodes)
    cands = oracle.knn_bruteforce(src_nodes.reshape(-1, dim), uniq, k) // P
    elem, xi, _, _ = oracle.locate(order, dim, src_nodes, uniq, cands.astype(np.int32), oracle.V1())
    vals = oracle.interp(order, dim, src_data, elem, xi)
    out = vals[recon].reshape(tgt_nodes.shape[0], tgt_nodes.shape[1], src_data.shape[1]).swapaxes(1, 2)
    return out, elem, xi


def test_gll_2_gll_config1_2d_quads(cuda, oracle, tmp_path):
    """BASELINE config 1: 2-D quad mesh, order-2 GLL -> second 2-D mesh, VP/VS/RHO, via the api."""
    import multi_mesh.api as api

    names = ["RHO", "VP", "VS"]
    src = meshgen.box_mesh((24, 24), 2, warp=0.01)
    tgt = meshgen.box_mesh((20, 20), 2, lo=[0.01, 0.02], hi=[0.97, 0.99])
    a, b = str(tmp_path / "from.npz"), str(tmp_path / "to.npz")
    src_data = _write(a, src, names)
    _write(b, tgt, names)
    api.gll_2_gll(a, b, nelem_to_search=20, parameters="ISO")
    with open_store(b, "r") as st:
        got = st.read("MODEL/data")
        assert st.labels("MODEL/data") == names
    want, _, _ = _oracle_gll_2_gll(oracle, src, src_data, tgt, 20)
    assert got.shape == want.shape == (400, 3, 9)
    assert np.max(np.abs(got - want) / np.abs(want)) <= 1e-10
    assert np.array_equal(got, want)
    exact = meshgen.analytic_fields(tgt, names, scale=None)  # smooth fields: interpolation error is small
    assert np.max(np.abs(got - want)) == 0 and np.isfinite(exact).all()


def test_gll_2_gll_3d_stored_array_and_fluid_fixup(cuda, oracle, tmp_path):
    import multi_mesh.api as api

    src = meshgen.box_mesh((6, 6, 6), 2)
    tgt = meshgen.box_mesh((5, 5, 5), 2, lo=[0.02] * 3, hi=[0.98] * 3)
    a, b, store = str(tmp_path / "from.npz"), str(tmp_path / "to.npz"), str(tmp_path / "stored")
    src_data = _write(a, src, ISO)
    fluid = np.zeros(125)
    fluid[:10] = 1
    old = _write(b, tgt, ISO, fluid=fluid)
    api.gll_2_gll(a, b, stored_array=store)
    want, elem, xi = _oracle_gll_2_gll(oracle, src, src_data, tgt, 20)
    want[:10] = old[:10]  # fluid elements keep their values (interpolator.py:829-830)
    with open_store(b, "r") as st:
        got = st.read("MODEL/data")
    assert np.array_equal(got, want)
    # stored interpolation matrices in the reference's format (interpolator.py:797-810)
    el = np.load(os.path.join(store, "elements.npy"), allow_pickle=True)
    co = np.load(os.path.join(store, "coeffs.npy"), allow_pickle=True)
    assert np.array_equal(el, elem)
    assert co.shape == (5, 27, len(elem)) and np.array_equal(co[0], co[4])
    assert np.array_equal(co[0].T, oracle.coeffs(2, 3, elem, xi))
    # second run (gradient=True: no fix-ups) takes the cached-matrix path
    _write(b, tgt, ISO, fluid=fluid)
    api.gll_2_gll(a, b, stored_array=store, gradient=True)
    full, _, _ = _oracle_gll_2_gll(oracle, src, src_data, tgt, 20)
    with open_store(b, "r") as st:
        got2 = st.read("MODEL/data")
    assert np.max(np.abs(got2 - full) / np.abs(full)) <= 1e-10


def _shell_pair(order):
    layers = meshgen.default_shell_layers()
    names = ISO + ["z_node_1D"]

    def build(n_lat):
        coords, el, z1d = meshgen.shell_mesh(n_lat, layers, order)
        data = meshgen.analytic_fields(coords, names)
        data[:, 5, :] = z1d
        ed = np.stack([el["fluid"], el["layer"]], axis=1)
        return SalvusMesh.from_arrays(coords, data, names, ed, ["fluid", "layer"], {"moho_idx": "2"})

    return build(4), build(3)


@pytest.mark.parametrize("order,driver", [(2, "layered"), (4, "layered"), (2, "multi"), (2, "multi_two"),
                                          (2, "points_layered")])
def test_layered_drivers(cuda, oracle, tmp_path, order, driver):
    """BASELINE config 3 (scaled down): cubed-sphere shell with a thin crust, curved order-n
    geometry, candidates restricted to the same layer (layer-local element ids)."""
    import multi_mesh.api as api
    from multi_mesh.components import interpolator as itp

    src, tgt = _shell_pair(order)
    layers = [1, 2, 3]
    src_fields = {p: src.element_nodal_fields[p].copy() for p in ISO}
    if driver == "layered":
        api.gll_2_gll_layered(src, tgt, layers=layers, parameters="ISO", stored_array=str(tmp_path / "st"))
        prm, k = oracle.V1(), 20
    elif driver == "multi":
        api.gll_2_gll_layered_multi(src, tgt, layers=layers, parameters=ISO, threads=4)
        prm, k = oracle.V1(), 20
    elif driver == "multi_two":
        api.gll_2_gll_layered_multi_two(src, tgt, layers=layers, parameters=ISO)
        prm, k = oracle.V2(1.05, True), 30
    else:
        itp.interpolate_to_points_layered(src, tgt, ISO, layers=layers)
        prm, k = oracle.V3(), 20
    masks_s, _ = utils.create_layer_mask(src, layers)
    uniq, masks_t, _ = utils.get_unique_points(tgt, mesh=True, layers=layers)
    for layer in map(str, layers):
        nodes_l = src.points[masks_s[layer]]
        pts, inv = uniq[layer]
        cands = oracle.knn_bruteforce(oracle.centroids(nodes_l), pts, k)
        elem, xi, _, nf = oracle.locate(order, 3, nodes_l, pts, cands, prm)
        fields_l = np.stack([src_fields[p][masks_s[layer]] for p in ISO], axis=1)
        vals = oracle.interp(order, 3, np.ascontiguousarray(fields_l), elem, xi)
        for f, p in enumerate(ISO):
            want = vals[inv, f].reshape(-1, tgt.n_gll_points)
            got = tgt.element_nodal_fields[p][masks_t[layer]]
            assert np.max(np.abs(got - want) / np.abs(want).max()) <= 1e-10, (layer, p)
            assert np.array_equal(got, want), (layer, p)
        if driver == "layered":
            with open_store(itp._interp_info_path(str(tmp_path / "st")), "r") as st:
                assert np.array_equal(st.read(f"elements/{layer}"), elem)
                assert np.array_equal(st.read(f"coeffs/{layer}"), oracle.coeffs(order, 3, elem, xi))
    if driver == "layered":  # cached matrices: second run reproduces the fields
        first = {p: tgt.element_nodal_fields[p].copy() for p in ISO}
        api.gll_2_gll_layered(src, tgt, layers=layers, parameters="ISO", stored_array=str(tmp_path / "st"))
        for p in ISO:
            assert np.max(np.abs(tgt.element_nodal_fields[p] - first[p]) / np.abs(first[p]).max()) <= 1e-10


def test_interpolate_to_points_and_query_model(cuda, oracle, tmp_path):
    import multi_mesh.api as api

    nodes = meshgen.box_mesh((5, 5, 5), 2, lo=[6.0e6, -2e5, -2e5], hi=[6.371e6, 2e5, 2e5], warp=0.01)
    data = meshgen.analytic_fields(nodes, ISO)
    mesh = SalvusMesh.from_arrays(nodes, data, ISO)
    rng = np.random.default_rng(1)
    pts = rng.uniform([5.95e6, -2.2e5, -2.2e5], [6.4e6, 2.2e5, 2.2e5], size=(3000, 3))
    got = api.interpolate_to_points(mesh, pts, ["VP", "RHO"])
    cands = oracle.knn_bruteforce(oracle.centroids(nodes), pts, 25)
    elem, xi, _, nf = oracle.locate(2, 3, nodes, pts, cands, oracle.V2())
    want = oracle.interp(2, 3, np.ascontiguousarray(data[:, [3, 2], :]), elem, xi)
    assert nf > 0 and np.array_equal(got, want) and (got[elem < 0] == 0).all()
    # query_model: lat/lon/depth -> xyz -> GLL-point k-NN -> V1
    path = str(tmp_path / "m.npz")
    write_gll_model(path, nodes, data, ISO)
    lld = np.stack([rng.uniform(-1.5, 1.5, 200), rng.uniform(-1.5, 1.5, 200), rng.uniform(1e3, 3.5e5, 200)], axis=1)
    vals = api.query_model(lld, path)
    xyz = utils.latlondepth_to_xyz(lld)
    c2 = oracle.knn_bruteforce(nodes.reshape(-1, 3), xyz, 20) // 27
    e2, x2, _, _ = oracle.locate(2, 3, nodes, xyz, c2.astype(np.int32), oracle.V1())
    assert np.array_equal(vals, oracle.interp(2, 3, data, e2, x2))


def test_exodus_round_trip_config4(cuda, oracle, tmp_path):
    """BASELINE config 4 (scaled down): exodus_2_gll then gll_2_exodus on an order-4 GLL mesh with
    gradient fields."""
    import multi_mesh.api as api

    points, conn = meshgen.hex8_mesh((10, 10, 10), warp=0.01)
    names = ["gradVP", "gradVS", "RHO"]
    lin = 2.0 + points[:, 0] + 2 * points[:, 1] + 3 * points[:, 2]
    nodal = {"gradVP": lin, "gradVS": np.sin(points[:, 0]) * np.cos(points[:, 1]), "RHO": 2600 + 300 * points[:, 2] ** 2}
    ex = Exodus.from_arrays(points, conn, nodal)
    gll_nodes = meshgen.box_mesh((4, 4, 4), 4, lo=[0.03] * 3, hi=[0.96] * 3)
    path = str(tmp_path / "gll.npz")
    write_gll_model(path, gll_nodes, np.zeros((64, 3, 125)), names)
    api.exodus_2_gll(ex, path, gll_order=4, parameters=names)
    with open_store(path, "r") as st:
        got = st.read("MODEL/data")
    # oracle: centroid k-NN, C-compat trilinear routine, nodal gather (interpolator.py:177-224)
    connC = np.ascontiguousarray(conn[:, np.argsort([0, 3, 2, 1, 4, 5, 6, 7])])
    q = gll_nodes.reshape(-1, 3)
    nn = oracle.knn_bruteforce(oracle.centroid_conn(conn, points), q, 20).astype(np.int64)
    nf, enc, w = oracle.trilinear_interpolator(20, nn, connC, points, q)
    assert nf == 0
    param = np.stack([nodal[n] for n in names])
    want = np.sum(param[:, enc] * w, axis=2).reshape(3, 64, 125).swapaxes(0, 1)
    assert np.max(np.abs(got - want)) <= 1e-10 * np.abs(want).max()
    assert np.max(np.abs(got[:, 0, :].reshape(-1) - (2.0 + q[:, 0] + 2 * q[:, 1] + 3 * q[:, 2]))) < 2e-3  # warped mesh
    # back: GLL -> exodus nodes that lie inside the GLL mesh
    inside = np.all((points > 0.05) & (points < 0.94), axis=1)
    ex2 = Exodus.from_arrays(points[inside], np.zeros((0, 8), dtype=np.int64), {n: np.zeros(inside.sum()) for n in names})
    api.gll_2_exodus(path, ex2, gll_order=4)
    back = ex2.get_nodal_field("gradVP")
    cands = oracle.knn_bruteforce(oracle.centroids(gll_nodes), points[inside], 20)
    e, x, _, _ = oracle.locate(4, 3, gll_nodes, points[inside], cands, oracle.V1())
    assert np.array_equal(back, oracle.interp(4, 3, got, e, x)[:, 0])
    assert np.max(np.abs(back - lin[inside])) < 5e-3  # round-trip error of the linear gradient field


def test_inner_operator_layouts(cuda, oracle):
    from multi_mesh.components import interpolator as itp
    from multimesh_b200.kdtree import KDTree

    rng = np.random.default_rng(4)
    nodes = meshgen.box_mesh((4, 4, 4), 2, warp=0.02)
    pts = rng.uniform(0, 1, (257, 3))
    tree = KDTree(nodes.reshape(-1, 3))
    dist, idx = tree.query(pts, k=20)
    o_idx, o_d2 = oracle.knn_bruteforce(nodes.reshape(-1, 3), pts, 20, return_d2=True)
    assert np.array_equal(idx, o_idx) and np.array_equal(dist, np.sqrt(o_d2))
    nearest = np.floor(idx / 27).astype(int)
    # find_gll_coeffs with the reference's transposed layouts (interpolator.py:758-782)
    coeffs = np.zeros((5, 27, len(pts)))
    element = np.zeros(len(pts))
    element, coeffs = itp.find_gll_coeffs(nodes, pts.T.copy(), nearest.T.copy(), coeffs, element, 3, 2, True)
    o_e, o_x, _, _ = oracle.locate(2, 3, nodes, pts, nearest.astype(np.int32), oracle.V1())
    assert np.array_equal(element, o_e) and np.array_equal(coeffs[0].T, oracle.coeffs(2, 3, o_e, o_x))
    assert not coeffs[1:].any()
    # scalar helpers
    e0 = int(o_e[0])
    assert np.array_equal(itp.inverse_transform(pts[0], nodes[e0], 3), o_x[0])
    assert np.array_equal(itp.get_coefficients(2, 2, 2, o_x[0], 3), oracle.weights(2, 3, o_x[0]))
    el, ref = itp._check_if_inside_element(nodes, nearest[0], pts[0], 3, True)
    assert el == e0 and np.array_equal(ref, o_x[0])
    assert itp.boundary_box_check(pts[0], nodes[e0])[0] is True
    assert np.isnan(itp.inverse_transform(np.array([50.0, 60.0, -70.0]), meshgen.box_mesh((2, 2, 2), 4, warp=0.08)[0], 3)).all()
    assert np.array_equal(itp._find_gll_centroids(nodes, 3), oracle.centroids(nodes))
    # get_element_weights (V2) with a centroid tree
    ctree = KDTree(oracle.centroids(nodes))
    elems, co = itp.get_element_weights(nodes, 2, ctree, pts, nelem_to_search=25, tolerance=1.05)
    c25 = oracle.knn_bruteforce(oracle.centroids(nodes), pts, 25)
    o_e2, o_x2, _, _ = oracle.locate(2, 3, nodes, pts, c25, oracle.V2())
    assert np.array_equal(elems, o_e2) and np.array_equal(co, oracle.coeffs(2, 3, o_e2, o_x2))


def test_map_to_sphere(cuda):
    from multi_mesh.components import interpolator as itp

    coords, el, z1d = meshgen.shell_mesh(2, meshgen.default_shell_layers(), 2)
    squash = coords * np.array([1.0, 1.0, 0.99])
    data = np.zeros((coords.shape[0], 1, 27))
    data[:, 0, :] = z1d
    m = SalvusMesh.from_arrays(squash.copy(), data, ["z_node_1D"])
    itp.map_to_sphere(m)
    x, y, z = squash[..., 0], squash[..., 1], squash[..., 2]
    r = np.sqrt(x ** 2 + y ** 2 + z ** 2)
    want = np.stack([x * 6371000 * z1d / r, y * 6371000 * z1d / r, z * 6371000 * z1d / r], axis=-1)
    assert np.array_equal(m.points, want)


def test_interpolate_to_mesh_and_plotter_values(cuda, oracle):
    """api.interpolate_to_mesh (api.py:353-393): both meshes mapped to the sphere, old -> new at the new mesh's
    nodes (V2, k = 25), coordinates restored; and the arrays the plotter draws (depth slice, cross section)."""
    import multi_mesh.api as api
    from multi_mesh.components import plotter

    names = ["VSV", "VSH", "VPV", "VPH", "z_node_1D"]

    def shell(n_lat, squash):
        coords, el, z1d = meshgen.shell_mesh(n_lat, meshgen.default_shell_layers(), 2)
        ell = coords * np.array([1.0, 1.0, squash])  # an "elliptic" mesh; z_node_1D keeps the spherical radius
        data = meshgen.analytic_fields(coords / 6371000.0, names[:4])
        data = np.concatenate([data, z1d[:, None, :]], axis=1)
        return coords, ell, SalvusMesh.from_arrays(ell.copy(), data.copy(), names), data

    sph_old, ell_old, old, data_old = shell(4, 0.995)
    sph_new, ell_new, new, _ = shell(3, 0.995)
    out = api.interpolate_to_mesh(old, new, params_to_interp=["VSV", "VPH"])
    assert out is new
    assert np.array_equal(old.points, ell_old) and np.array_equal(new.points, ell_new)  # coordinates restored
    # oracle on the sphere-mapped meshes (map_to_sphere is bit-equal to the numpy formula, test_map_to_sphere)
    def to_sphere(ell, z1d):  # the operation order of interpolator.py:1138-1143, as in test_map_to_sphere
        x, y, z = ell[..., 0], ell[..., 1], ell[..., 2]
        r = np.sqrt(x ** 2 + y ** 2 + z ** 2)
        return np.stack([x * 6371000 * z1d / r, y * 6371000 * z1d / r, z * 6371000 * z1d / r], axis=-1)

    s_old = to_sphere(ell_old, data_old[:, 4, :])
    s_new = to_sphere(ell_new, new.element_nodal_fields["z_node_1D"])
    pts = s_new.reshape(-1, 3)
    cands = oracle.knn_bruteforce(oracle.centroids(s_old), pts, 25)
    elem, xi, _, _ = oracle.locate(2, 3, s_old, pts, cands, oracle.V2())
    want = oracle.interp(2, 3, np.ascontiguousarray(data_old[:, [0, 3], :]), elem, xi)
    assert np.array_equal(new.element_nodal_fields["VSV"], want[:, 0].reshape(new.nelem, 27))
    assert np.array_equal(new.element_nodal_fields["VPH"], want[:, 1].reshape(new.nelem, 27))
    # depth slice through a spherical mesh: the generator's cloud -> latlondepth_to_xyz -> interpolate_to_points
    sphere_mesh = SalvusMesh.from_arrays(sph_old.copy(), data_old.copy(), names)
    vals = plotter.depth_slice_values(sphere_mesh, 300.0, 12, "VSV", lat_extent=(-40.0, 40.0), lon_extent=(-60.0, 60.0))
    cloud = utils.latlondepth_to_xyz(plotter._create_depthslice(300e3, 12, (-40.0, 40.0), (-60.0, 60.0)))
    c2 = oracle.knn_bruteforce(oracle.centroids(sph_old), cloud, 25)
    e2, x2, _, _ = oracle.locate(2, 3, sph_old, cloud, c2, oracle.V2())
    assert np.array_equal(vals, oracle.interp(2, 3, np.ascontiguousarray(data_old[:, :1, :]), e2, x2).reshape(12, 12))
    sec = plotter.cross_section_values(sphere_mesh, 10.0, 20.0, -30.0, 80.0, "VSV", npoints=9, nrads=5,
                                       min_depth_in_km=50.0, max_depth_in_km=2000.0, relative=False)
    assert sec.shape == (5, 9) and np.isfinite(sec).all() and (sec != 0).any()
