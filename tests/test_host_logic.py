"""CPU tests of the host-side logic: layer handling, unique points, stores, readers, partitioning."""
import os

import numpy as np
import pytest

from multimesh_b200 import meshgen, utils
from multimesh_b200.components.salvus_mesh_reader import SalvusMesh
from multimesh_b200.io.store import open_store, write_gll_model


def _shell(order=2, n_lat=2):
    layers = [(3480e3, 5000e3, 2, 4, 1), (5000e3, 6291e3, 2, 3, 0), (6291e3, 6346e3, 1, 2, 0), (6346e3, 6371e3, 1, 1, 0)]
    coords, elemental, z1d = meshgen.shell_mesh(n_lat, layers, order)
    names = ["VP", "VS", "z_node_1D"]
    data = meshgen.analytic_fields(coords, names)
    data[:, 2, :] = z1d
    ed = np.stack([elemental["layer"], elemental["fluid"]], axis=1)
    return SalvusMesh.from_arrays(coords, data, names, ed, ["layer", "fluid"], {"moho_idx": "2"})


def test_pick_parameters():
    assert utils.pick_parameters("ISO") == ["QKAPPA", "QMU", "RHO", "VP", "VS"]
    assert utils.pick_parameters("TTI")[:2] == ["VPV", "VPH"] and len(utils.pick_parameters("TTI")) == 8
    assert utils.pick_parameters(["A"]) == ["A"]


def test_assess_layers_and_masks():
    m = _shell()
    layers, mask = utils._assess_layers(m, "all")
    assert list(layers) == [4, 3, 2, 1] and mask is False
    assert list(utils._assess_layers(m, "crust")[0]) == [4, 3]  # first moho_idx layers of the descending list
    assert list(utils._assess_layers(m, "nocore")[0]) == []  # fluid layer 4 is first in the descending list
    assert list(utils._assess_layers(m, "core")[0]) == [4, 3, 2, 1]
    assert utils._assess_layers(m, [1, 2]) == ([1, 2], True)
    assert utils._assess_layers(m, 3) == ([3], True)
    with pytest.raises(ValueError):
        utils._assess_layers(m, [9])
    with pytest.raises(ValueError):
        utils._assess_layers(m, "nonsense")
    masks, layers = utils.create_layer_mask(m, [1, 3])
    assert set(masks) == {"1", "3"}
    assert masks["1"].sum() + masks["3"].sum() == (m.elemental_fields["layer"] == 1).sum() + (m.elemental_fields["layer"] == 3).sum()
    assert not (masks["1"] & masks["3"]).any()


@pytest.mark.gpu
def test_get_unique_points_array_and_mesh():
    """utils.get_unique_points runs on the GPU (K4); the array form must equal np.unique bit for bit."""
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    nodes = meshgen.box_mesh((3, 3, 3), 2)
    u, inv = utils.get_unique_points(nodes)
    nu, ninv = np.unique(nodes.reshape(-1, 3), return_inverse=True, axis=0)
    assert np.array_equal(u, nu) and np.array_equal(inv, ninv.reshape(-1))
    assert u.shape == (7 ** 3, 3)
    assert np.array_equal(u[inv].reshape(nodes.shape), nodes)
    assert (np.diff(u.view([("x", "f8"), ("y", "f8"), ("z", "f8")]).ravel().argsort(order=("x", "y", "z"))) == 1).all()
    m = _shell()
    up, mask, layers = utils.get_unique_points(m, mesh=True, layers=[1, 2])
    for k in ("1", "2"):
        pts, inv = up[k]
        assert np.array_equal(pts[inv].reshape(-1, m.n_gll_points, 3), m.points[mask[k]])


def test_latlondepth_to_xyz():
    xyz = utils.latlondepth_to_xyz(np.array([[90.0, 0.0, 0.0], [0.0, 0.0, 1000.0], [0.0, 90.0, 0.0]]))
    assert np.allclose(xyz[0], [0, 0, 6371000.0], atol=1e-6)
    assert np.allclose(xyz[1], [6370000.0, 0, 0], atol=1e-6)
    assert np.allclose(xyz[2], [0, 6371000.0, 0], atol=1e-6)


def test_npz_store_and_salvus_mesh_round_trip(tmp_path):
    nodes = meshgen.box_mesh((2, 2, 2), 2)
    names = ["QKAPPA", "QMU", "RHO", "VP", "VS"]
    data = meshgen.analytic_fields(nodes, names)
    ed = np.zeros((8, 2))
    path = str(tmp_path / "model.npz")
    write_gll_model(path, nodes, data, names, ed, ["fluid", "layer"], {"moho_idx": "1"})
    pts, dat, params = utils.load_hdf5_params_to_memory(path, "MODEL/data", "MODEL/coordinates")
    assert np.array_equal(pts, nodes) and np.array_equal(dat, data) and params == names
    m = SalvusMesh(path, fast_mode=False)
    assert (m.nelem, m.n_gll_points, m.dimensions, m.shape_order) == (8, 27, 3, 2)
    assert m.global_strings["moho_idx"] == "1"
    assert np.array_equal(m.get_element_centroids(), nodes.mean(axis=1))
    new = np.full((8, 27), 7.0)
    m.attach_field("VP", new)
    assert np.array_equal(SalvusMesh(path, fast_mode=False).element_nodal_fields["VP"], new)
    with pytest.raises(ValueError):
        m.attach_field("NOPE", new)
    with pytest.raises(ValueError):
        m.attach_field("VP", np.zeros((3, 3)))
    with open_store(path, "r+") as st:
        utils.remove_and_create_empty_dataset(st, ["A", "B"], "MODEL/data", "MODEL/coordinates")
    with open_store(path, "r") as st:
        assert st.shape("MODEL/data") == (8, 2, 27) and st.labels("MODEL/data") == ["A", "B"]


def test_grad_labels_are_stripped(tmp_path):
    nodes = meshgen.box_mesh((2, 2, 2), 1)
    path = str(tmp_path / "g.npz")
    write_gll_model(path, nodes, np.zeros((8, 2, 8)), ["gradVP", "gradVS"])
    assert utils.load_hdf5_params_to_memory(path, "MODEL/data", "MODEL/coordinates")[2] == ["VP", "VS"]


def test_shard_bounds():
    from multimesh_b200.parallel import local_slice, shard_bounds

    for n, w in [(10, 3), (7, 8), (0, 2), (100, 1), (23887872, 8)]:
        b = shard_bounds(n, w)
        assert b[0] == 0 and b[-1] == n and len(b) == w + 1
        sizes = np.diff(b)
        assert sizes.max() - sizes.min() <= 1
        assert sum(local_slice(n, r, w).stop - local_slice(n, r, w).start for r in range(w)) == n


def test_slab_partition():
    """Equal-count slabs along the longest axis; a partition of all indices; deterministic; edge cases."""
    from multimesh_b200.parallel import slab_partition

    rng = np.random.default_rng(0)
    pts = rng.random((10_001, 3)) * np.array([1.0, 5.0, 2.0])  # longest extent: y
    for world in (1, 2, 3, 8):
        sets = slab_partition(pts, world)
        assert len(sets) == world
        allidx = np.concatenate(sets)
        assert np.array_equal(np.sort(allidx), np.arange(len(pts)))
        sizes = [len(s) for s in sets]
        assert max(sizes) - min(sizes) <= 1
        for a, b in zip(sets[:-1], sets[1:]):
            assert pts[a, 1].max() <= pts[b, 1].min()
        again = slab_partition(pts.copy(), world)
        assert all(np.array_equal(a, b) for a, b in zip(sets, again))
    # explicit axis, more ranks than points, no points, ties on the cut stay together
    assert pts[slab_partition(pts, 2, axis=0)[0], 0].max() <= pts[slab_partition(pts, 2, axis=0)[1], 0].min()
    few = slab_partition(pts[:3], 8)
    assert sum(len(s) for s in few) == 3 and len(few) == 8
    assert [len(s) for s in slab_partition(pts[:0], 4)] == [0, 0, 0, 0]
    tied = np.zeros((100, 3))
    tied[:, 0] = np.repeat([0.0, 1.0], 50)
    t = slab_partition(tied, 2)
    assert np.array_equal(np.sort(np.concatenate(t)), np.arange(100)) and {len(t[0]), len(t[1])} == {50}


def test_meshgen_shapes_and_shell_radii():
    c = meshgen.box_mesh((3, 2), 4, lo=[0, 0], hi=[3, 2])
    assert c.shape == (6, 25, 2) and c.min() == 0 and c[..., 0].max() == 3
    coords, el, z1d = meshgen.shell_mesh(2, meshgen.default_shell_layers(), 2)
    r = np.linalg.norm(coords, axis=2)
    assert np.allclose(r, z1d * 6371000.0, rtol=1e-13)
    assert r.min() >= 3480e3 - 1 and r.max() <= 6371e3 + 1
    assert set(np.unique(el["layer"])) == {1.0, 2.0, 3.0}
    pts, conn = meshgen.hex8_mesh((2, 3, 4))
    assert pts.shape == (3 * 4 * 5, 3) and conn.shape == (24, 8) and conn.max() == len(pts) - 1


def test_plotter_point_generators_and_geodesy():
    """Point-cloud generators of the plotter (components/plotter.py:159-187, 361-380) and their geodesy helpers."""
    from multi_mesh.components import plotter

    pts = plotter._create_depthslice(150e3, 5, lat_extent=(-10.0, 10.0), lon_extent=(20.0, 40.0))
    lat, lon = np.linspace(-10, 10, 5), np.linspace(20, 40, 5)
    xx, yy = np.meshgrid(lat, lon)
    assert pts.shape == (25, 3) and np.array_equal(pts[:, 0], xx.ravel()) and np.array_equal(pts[:, 1], yy.ravel())
    assert (pts[:, 2] == 150e3).all()
    gc = utils.greatcircle_points(10.0, 20.0, -30.0, 80.0, npts=50)
    assert gc.shape == (50, 2) and np.allclose(gc[0], [10.0, 20.0])
    v = np.stack([np.cos(np.deg2rad(gc[:, 0])) * np.cos(np.deg2rad(gc[:, 1])),
                  np.cos(np.deg2rad(gc[:, 0])) * np.sin(np.deg2rad(gc[:, 1])), np.sin(np.deg2rad(gc[:, 0]))], axis=1)
    step = np.arccos(np.clip((v[1:] * v[:-1]).sum(axis=1), -1, 1))
    assert np.allclose(step, step[0], rtol=1e-9)          # equally spaced ...
    normal = np.cross(v[0], v[-1])
    assert np.allclose(v @ normal, 0.0, atol=1e-12)       # ... on one great circle
    with pytest.raises(Exception):
        utils.greatcircle_points(0, 0, 1, 1, npts=2)
    assert utils.elliptic_to_geocentric_latitude(0.0) == 0.0 and utils.elliptic_to_geocentric_latitude(90.0) == 90.0
    assert abs(utils.elliptic_to_geocentric_latitude(45.0) - 44.80757678) < 1e-6   # known WGS84 value
    x, y, z = utils.sph2cart(np.array([0.3]), np.array([1.1]), np.array([6.0e6]))
    c, l, r = utils.cart2sph(x, y, z)
    assert np.allclose([c[0], l[0], r[0]], [0.3, 1.1, 6.0e6])
    with pytest.raises(ValueError):
        utils.sph2cart(np.array([-0.1]), np.array([0.0]), np.array([1.0]))
    xyz, rads = plotter.cross_section_points(10.0, 20.0, -30.0, 80.0, npoints=7, nrads=4, min_depth_in_km=0.0,
                                             max_depth_in_km=600.0)
    assert xyz.shape == (28, 3) and np.allclose(np.linalg.norm(xyz, axis=1).reshape(4, 7), rads[:, None])


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU port of the path on the host cores; the one place outside tests/ that may
    execute oracle/) prints ONE JSON line with the contract's keys, the same `config` object our arm prints, and a
    cpu_baseline / e2e description of the run -- checked on the small workload, no GPU involved."""
    import json
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--workload", "small",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=300, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "impl"):
        assert key in d, key
    assert d["impl"] == "reference" and d["value"] > 0 and d["unit"] == "points/s" and d["nfailed"] == 0
    assert set(d["config"]) == {"workload", "points_per_gpu", "source_elements", "fields", "l2"}
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0


def test_bench_launch_count_matches_the_committed_launch_list():
    """bench.py's `gpu_launches` claim per step == the kernels of one fused step in the committed ncu launch list."""
    import bench

    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles",
                        "r2_S2_fused_step_launches.txt")
    lines = [ln for ln in open(path).read().splitlines() if ln.strip().endswith(" us")]
    kernels = [ln for ln in lines if not ln.startswith("fused step total")]
    assert len(kernels) == bench.launches_per_step(23_887_872, True, True) == 29
    assert bench.launches_per_step(100_000_000, True, False) == 34  # centroid form: + grouping by first candidate
