"""
Generates tests/golden/glue_*.npz by RUNNING THE REFERENCE'S OWN PYTHON (imported from
/root/reference through tests/golden/refglue.py) on small seeded meshes:

  glue_v1_*.npz      find_gll_coeffs -> _check_if_inside_element            interpolator.py:1540-1597, 1409-1473
  glue_v2_*.npz      get_element_weights (tolerance, snap_to_nearest)       interpolator.py:1147-1255
  glue_v3.npz        get_element_weights_layered                            interpolator.py:1258-1334
  glue_v4.npz        v2_interpolation_tools.get_element_weights             v2_interpolation_tools.py:71-164
  glue_v5.npz        scripts/cli.py _check_if_inside_element                cli.py:401-430
  glue_points.npz    interpolate_to_points (centroid tree, V2, gather)      interpolator.py:931-977
  glue_gll2gll_2d.npz  the same driver on 2-D order-4 quads (BASELINE config 1's shape)
  glue_gll2gll.npz   gll_2_gll end to end (dedup, GLL-point tree // P, V1, gather, recon, fluid
                     fix-up) on in-memory files                             interpolator.py:621-852

What these pin: the reference's control flow (candidate order, accept predicates, AABB prefilter,
fall-backs and their tie-breaks, -1 / zero-weight handling, de-duplication + reconstruction, the
gather, the fluid/solid fix-up).  What they do NOT pin: the arithmetic inside salvus.fem (closed
source) and pykdtree's tie order -- both are served by the oracle here, see refglue.py.

Run in the build container:  OMP_NUM_THREADS=1 python tests/golden/make_golden_glue.py
"""
import os
import sys

os.environ.setdefault("OMP_NUM_THREADS", "1")  # the reference forks worker pools; keep libgomp single-threaded
sys.dont_write_bytecode = True  # /root/reference is read-only

import numpy as np  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import refglue  # noqa: E402

interp, v2, ref_utils = refglue.load_reference()
cli = refglue.load_reference_cli()

from multimesh_b200 import meshgen  # noqa: E402  (after the reference: `multi_mesh` must be the reference's)
from oracle import capi  # noqa: E402
import glue_inputs  # noqa: E402

capi.set_num_threads(1)
NAMES = ["QKAPPA", "QMU", "RHO", "VP", "VS"]
SEEDS = {"case_v3": 301, "case_v4": 401, "case_v5": 501, "case_points": 601, "case_gll2gll": 701, "case_layered": 801,
         "case_query_model": 901, "case_exodus": 1001,
         "case_gll2gll_2d": 1101}


def save(name, **kw):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **kw)
    print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB")


def targets(rng, nodes, n_random, stride, lo=-0.08, hi=1.08):
    d = nodes.shape[2]
    return np.concatenate([rng.uniform(lo, hi, (n_random, d)), nodes.reshape(-1, d)[::stride]])


def run_v1(nodes, pts, cands):
    """find_gll_coeffs exactly as gll_2_gll calls it (:790-799): [d,N] points, [k,N] candidates."""
    N, P, d = len(pts), nodes.shape[1], nodes.shape[2]
    order = round(P ** (1.0 / d)) - 1
    element, coeffs = interp.find_gll_coeffs(
        original_coordinates=nodes, coordinates=np.swapaxes(pts, 0, 1), nearest_elements=np.swapaxes(cands, 0, 1),
        coeffs=np.zeros((1, P, N)), element=np.zeros(N), dimensions=d, from_gll_order=order,
        ignore_hard_elements=True)
    return element.astype(np.int64), np.ascontiguousarray(coeffs[0].T)


def case_v1(tag, shape, order, warp, n_random, stride, seed):
    rng = np.random.default_rng(seed)
    nodes = meshgen.box_mesh(shape, order, warp=warp)
    E, P, d = nodes.shape
    pts = targets(rng, nodes, n_random, stride)
    k = min(20, E)
    cen = refglue.CanonicalKDTree(capi.centroids(nodes)).query(pts, k)[1]
    gll = np.floor(refglue.CanonicalKDTree(nodes.reshape(-1, d)).query(pts, k)[1] / P).astype(int)  # :751-756
    e1, c1 = run_v1(nodes, pts, cen)
    e2, c2 = run_v1(nodes, pts, gll)
    save(f"glue_v1_{tag}.npz", shape=np.array(shape), order=order, warp=warp, points=pts, k=k,
         cands_centroid=cen.astype(np.int32), cands_gll=gll.astype(np.int32), elem_centroid=e1, coeffs_centroid=c1,
         elem_gll=e2, coeffs_gll=c2)


def case_v2(tag, shape, order, warp, n_random, stride, seed, tolerance):
    rng = np.random.default_rng(seed)
    nodes = meshgen.box_mesh(shape, order, warp=warp)
    pts = targets(rng, nodes, n_random, stride, -0.15, 1.15)
    tree = refglue.CanonicalKDTree(capi.centroids(nodes))
    k = min(25, len(nodes))
    out = {}
    for snap in (False, True):
        e, c = interp.get_element_weights(nodes, order, tree, pts, nelem_to_search=k, tolerance=tolerance,
                                          snap_to_nearest=snap)
        out[f"elem_snap{int(snap)}"] = np.asarray(e, dtype=np.int64)
        out[f"coeffs_snap{int(snap)}"] = np.asarray(c, dtype=np.float64)
    save(f"glue_v2_{tag}.npz", shape=np.array(shape), order=order, warp=warp, points=pts, k=k, tolerance=tolerance,
         **out)


class DuckMesh:
    """The attributes the reference reads from a salvus UnstructuredMesh on these paths."""

    def __init__(self, nodes, fields, names):
        E, P, d = nodes.shape
        self.points_flat = nodes.reshape(E * P, d)
        self.connectivity = np.arange(E * P).reshape(E, P)
        self.shape_order = round(P ** (1.0 / d)) - 1
        self.element_nodal_fields = {n: fields[:, i, :] for i, n in enumerate(names)}
        self.n_gll_points = P
        self.points = self.points_flat

    def get_element_centroid(self):
        return np.mean(self.points[self.connectivity], axis=1)


def case_v3(seed):
    rng = np.random.default_rng(seed)
    nodes = meshgen.box_mesh((4, 4, 6), 2, warp=0.03)
    cz = capi.centroids(nodes)[:, 2]
    layer_of = np.where(cz < 0.34, 2, np.where(cz < 0.67, 1, 0))
    pts = targets(rng, nodes, 700, 9, -0.05, 1.05)
    new_coordinates, nearest, mask = {}, {}, {}
    for lay in (0, 1, 2):
        m = layer_of == lay
        lo, hi = {2: (-1, 0.34), 1: (0.30, 0.70), 0: (0.64, 2)}[lay]  # overlapping bands: some points miss their layer
        sel = (pts[:, 2] >= lo) & (pts[:, 2] <= hi)
        mask[str(lay)] = m
        new_coordinates[str(lay)] = (pts[sel],)
        nearest[str(lay)] = refglue.CanonicalKDTree(capi.centroids(nodes[m])).query(pts[sel], 20)[1]

    class M:  # get_element_weights_layered indexes original_mesh.points[original_mask[layer]] (:1277)
        points = nodes

    elems, coeffs = interp.get_element_weights_layered(new_coordinates, nearest, M, mask, dimensions=3,
                                                       from_gll_order=2)
    out = {"shape": np.array((4, 4, 6)), "order": 2, "warp": 0.03, "layer_of": layer_of}
    for lay in ("0", "1", "2"):
        out[f"points_{lay}"] = new_coordinates[lay][0]
        out[f"cands_{lay}"] = nearest[lay].astype(np.int32)
        out[f"elem_{lay}"] = np.asarray(elems[lay], dtype=np.int64)
        out[f"coeffs_{lay}"] = np.asarray(coeffs[lay], dtype=np.float64)
    save("glue_v3.npz", **out)


def case_v4(seed):
    rng = np.random.default_rng(seed)
    nodes = meshgen.box_mesh((5, 5, 4), 2, warp=0.03)
    pts = targets(rng, nodes, 800, 11, -0.12, 1.12)
    tree = refglue.CanonicalKDTree(capi.centroids(nodes))
    e, c = v2.get_element_weights(nodes, tree, pts)
    save("glue_v4.npz", shape=np.array((5, 5, 4)), order=2, warp=0.03, points=pts, k=25,
         elem=np.asarray(e, dtype=np.int64), coeffs=np.asarray(c, dtype=np.float64))


def case_v5(seed):
    import contextlib
    import io

    rng = np.random.default_rng(seed)
    nodes = meshgen.box_mesh((3, 3, 3), 4, warp=0.02)
    out = {}
    # k = 20: far candidates of outside points do not converge -> the reference accepts their NaN xi (cli.py:421,
    # `not any(nan > 1.02)`); k = 3: every candidate converges, so the min-sum|xi| fall-back (:424-428) is reached.
    for k, lo, hi in ((20, -0.1, 1.1), (3, -0.04, 1.04)):
        pts = targets(rng, nodes, 300, 31, lo, hi)
        cands = refglue.CanonicalKDTree(capi.centroids(nodes)).query(pts, k)[1]
        elem = np.zeros(len(pts), dtype=np.int64)
        xi = np.zeros((len(pts), 3))
        with contextlib.redirect_stdout(io.StringIO()):
            for i, p in enumerate(pts):
                elem[i], xi[i] = cli._check_if_inside_element(nodes, cands[i], p)
        out.update({f"points_k{k}": pts, f"cands_k{k}": cands.astype(np.int32), f"elem_k{k}": elem, f"xi_k{k}": xi})
    save("glue_v5.npz", shape=np.array((3, 3, 3)), order=4, warp=0.02, **out)


def case_points(seed):
    rng = np.random.default_rng(seed)
    nodes = meshgen.box_mesh((5, 4, 6), 2, warp=0.03)
    fields = meshgen.analytic_fields(nodes, NAMES)
    pts = targets(rng, nodes, 900, 13, -0.1, 1.1)
    vals = interp.interpolate_to_points(DuckMesh(nodes, fields, NAMES), pts, NAMES)
    save("glue_points.npz", shape=np.array((5, 4, 6)), order=2, warp=0.03, points=pts, names=np.array(NAMES),
         values=vals)


def case_gll2gll(seed):
    rng = np.random.default_rng(seed)
    src = meshgen.box_mesh((5, 4, 6), 2, warp=0.03)
    tgt = meshgen.box_mesh((4, 5, 5), 2, warp=0.02)
    fields = meshgen.analytic_fields(src, NAMES)
    # a block of the source with VS == 0 (a "fluid" region) so the fake-fluid fix-up (:829-841) has work to do
    cx = capi.centroids(src)[:, 0]
    fields[cx < 0.25, NAMES.index("VS"), :] = 0.0
    old = rng.uniform(1.0, 2.0, (tgt.shape[0], len(NAMES), tgt.shape[1]))
    fluid = (capi.centroids(tgt)[:, 2] > 0.85).astype(np.float64)
    edata = np.stack([fluid, np.arange(len(tgt), dtype=np.float64)], axis=1)
    refglue.write_gll_file("from.h5", src, fields, NAMES)
    refglue.write_gll_file("to.h5", tgt, old, NAMES, element_data=edata, element_labels=["fluid", "layer"])
    # interpolator.py:814 indexes with the float `element` array find_gll_coeffs hands back (IndexError on every
    # numpy >= 1.12 unless points failed and :803 cast it); apply that same cast so the driver can run.
    original = interp.find_gll_coeffs

    def find_gll_coeffs_int(**kw):
        element, coeffs = original(**kw)
        return element.astype(int), coeffs

    interp.find_gll_coeffs = find_gll_coeffs_int
    try:
        interp.gll_2_gll("from.h5", "to.h5", nelem_to_search=20)
    finally:
        interp.find_gll_coeffs = original
    out = refglue.FILES["to.h5"]["MODEL/data"].array
    label = refglue.FILES["to.h5"]["MODEL/data"].dims[1].label
    save("glue_gll2gll.npz", src_shape=np.array((5, 4, 6)), src_warp=0.03, tgt_shape=np.array((4, 5, 5)),
         tgt_warp=0.02, order=2, names=np.array(NAMES), source_fields=fields, target_old=old, fluid=fluid,
         values=out, label=np.array(label))


def case_gll2gll_2d(seed):
    """gll_2_gll on 2-D quads (BASELINE config 1; the reference's 2-D dispatch exists for order 4 only, :50-53)."""
    import contextlib
    import io

    rng = np.random.default_rng(seed)
    names = ["RHO", "VP", "VS"]
    src = meshgen.box_mesh((7, 6), 4, warp=0.02)
    tgt = meshgen.box_mesh((5, 6), 4, lo=[0.01, 0.02], hi=[0.98, 0.97], warp=0.01)
    fields = meshgen.analytic_fields(src, names)
    old = rng.uniform(1.0, 2.0, (tgt.shape[0], len(names), tgt.shape[1]))
    fluid = np.zeros(len(tgt))
    edata = np.stack([fluid, np.ones(len(tgt))], axis=1)
    refglue.write_gll_file("from2d.h5", src, fields, names)
    refglue.write_gll_file("to2d.h5", tgt, old, names, element_data=edata, element_labels=["fluid", "layer"])
    original = interp.find_gll_coeffs

    def find_gll_coeffs_int(**kw):  # same float-index bridge as case_gll2gll
        element, coeffs = original(**kw)
        return element.astype(int), coeffs

    interp.find_gll_coeffs = find_gll_coeffs_int
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            interp.gll_2_gll("from2d.h5", "to2d.h5", nelem_to_search=20)
    finally:
        interp.find_gll_coeffs = original
    save("glue_gll2gll_2d.npz", src_shape=np.array((7, 6)), src_warp=0.02, tgt_shape=np.array((5, 6)), tgt_warp=0.01,
         order=4, names=np.array(names), values=refglue.FILES["to2d.h5"]["MODEL/data"].array)


def shell_files(order, seed, layers=None):
    pair = glue_inputs.shell_pair(order, seed, layers)
    for key, (coords, data, ed) in pair.items():
        refglue.write_gll_file(f"shell_{key}.h5", coords, data, NAMES + ["z_node_1D"], element_data=ed,
                               element_labels=["fluid", "layer"], global_strings={"moho_idx": "2"})


STRIDE = {"layered": 1, "layered_o4": 5, "multi": 2, "multi_two": 2, "points_layered": 2, "multi_subset": 1,
          "multi_two_nocore": 1}
CORE_MESH = ("layered_o4", "multi_subset", "multi_two_nocore")  # cases run on the mesh that has a fluid core layer


def case_layered(seed):
    """gll_2_gll_layered (V1 per layer, :288-439), gll_2_gll_layered_multi (:442-618),
    gll_2_gll_layered_multi_two (V2 snap, k=30, :980-1082), interpolate_to_points_layered (V3, :855-928)."""
    import contextlib
    import io

    res = {}
    runs = [("layered", 2, lambda: interp.gll_2_gll_layered("shell_from.h5", "shell_to.h5", layers=[1, 2, 3],
                                                              parameters="ISO")),
            ("layered_o4", 4, lambda: interp.gll_2_gll_layered("shell_from.h5", "shell_to.h5", layers="nocore",
                                                                 parameters="ISO")),
            ("multi", 2, lambda: interp.gll_2_gll_layered_multi("shell_from.h5", "shell_to.h5", layers=[1, 2, 3],
                                                                 parameters=NAMES, threads=3)),
            ("multi_two", 2, lambda: interp.gll_2_gll_layered_multi_two("shell_from.h5", "shell_to.h5",
                                                                         layers=[1, 2, 3], parameters=NAMES)),
            ("points_layered", 2, lambda: interp.interpolate_to_points_layered("shell_from.h5", "shell_to.h5", NAMES,
                                                                                layers=[1, 2, 3])),
            # layer SUBSETS on the mesh with a core: the multi drivers start from the target's existing fields
            # (:607, :1070), so elements outside the requested layers must keep their values
            ("multi_subset", 2, lambda: interp.gll_2_gll_layered_multi("shell_from.h5", "shell_to.h5", layers=[2, 1],
                                                                        parameters=NAMES, threads=2)),
            ("multi_two_nocore", 2, lambda: interp.gll_2_gll_layered_multi_two("shell_from.h5", "shell_to.h5",
                                                                                layers="nocore", parameters=NAMES))]
    for tag, order, run in runs:
        shell_files(order, seed, glue_inputs.CORE_LAYERS if tag in CORE_MESH else None)
        with contextlib.redirect_stdout(io.StringIO()):
            run()
        vals = refglue.FILES["shell_to.h5"]["MODEL/data"].array[:, :5, :]
        res[f"{tag}_values"] = vals[::STRIDE[tag]].copy()  # every STRIDE-th target element (fixture size)
        res[f"{tag}_order"] = order
        res[f"{tag}_stride"] = STRIDE[tag]
    save("glue_layered.npz", seed=seed, **res)


def case_query_model(seed):
    rng = np.random.default_rng(seed)
    nodes = meshgen.box_mesh((5, 5, 5), 2, lo=[6.0e6, -2e5, -2e5], hi=[6.371e6, 2e5, 2e5], warp=0.01)
    data = meshgen.analytic_fields(nodes, NAMES)
    refglue.write_gll_file("model.h5", nodes, data, NAMES)
    lld = np.stack([rng.uniform(-1.5, 1.5, 300), rng.uniform(-1.5, 1.5, 300), rng.uniform(1e3, 3.5e5, 300)], axis=1)
    vals = interp.query_model(lld, "model.h5", 20, "MODEL/data", "MODEL/coordinates")
    save("glue_query_model.npz", latlondepth=lld, values=vals)


def case_exodus(seed):
    """exodus_2_gll (reference Python + the reference's own compiled C, :142-224) and gll_2_exodus (V1, :227-285)."""
    import contextlib
    import io

    from oracle import build as oracle_build

    oracle_build.build_ref()
    refglue.load_reference_exodus(interp)
    names = ["VP", "VS", "RHO"]
    points, conn = meshgen.hex8_mesh((8, 7, 6), warp=0.02)
    nodal = {"VP": 2.0 + points[:, 0] + 2 * points[:, 1] + 3 * points[:, 2],
             "VS": np.sin(points[:, 0]) * np.cos(points[:, 1]) + 2.0, "RHO": 2600 + 300 * points[:, 2] ** 2}
    refglue.write_exodus_file("mesh.e", points, conn, nodal)
    gll = meshgen.box_mesh((3, 3, 2), 4, lo=[0.04] * 3, hi=[0.95] * 3, warp=0.01)
    f = refglue.write_gll_file("gll.h5", gll, np.zeros((len(gll), 3, 125)), names)
    f["MODEL/data"].attrs["DIMENSION_LABELS"] = [b"element", ("[ " + " | ".join(names) + " ]").encode(), b"point"]
    with contextlib.redirect_stdout(io.StringIO()):
        interp.exodus_2_gll("mesh.e", "gll.h5", gll_order=4, parameters=names)
    on_gll = refglue.FILES["gll.h5"]["MODEL/data"].array.copy()
    # back onto exodus nodes inside the GLL mesh.  create_dimension_labels (utils.py:159-168) only sets the
    # dimension-scale labels; real h5py then serves them through attrs["DIMENSION_LABELS"], the fake does not:
    ds = refglue.FILES["gll.h5"]["MODEL/data"]
    ds.attrs["DIMENSION_LABELS"] = [b"element", ds.dims[1].label.encode(), b"point"]
    inside = np.all((points > 0.07) & (points < 0.92), axis=1)
    refglue.write_exodus_file("back.e", points[inside], np.zeros((0, 8), dtype=np.int64),
                              {n: np.zeros(int(inside.sum())) for n in names})
    # gll_2_exodus calls _check_if_inside_element with four arguments (:274-276) although it takes five (:1409-1411):
    # TypeError as shipped.  Supply the missing `ignore_hard_elements` with the value its other call sites use (True).
    original = interp._check_if_inside_element

    def with_default(gll_model, nearest_elements, point, dimension, ignore_hard_elements=True):
        return original(gll_model, nearest_elements, point, dimension, ignore_hard_elements)

    interp._check_if_inside_element = with_default
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            interp.gll_2_exodus("gll.h5", "back.e", gll_order=4)
    finally:
        interp._check_if_inside_element = original
    back = np.stack([refglue.EXO["back.e"]["nodal"][n] for n in names], axis=1)
    save("glue_exodus.npz", hex_shape=np.array((8, 7, 6)), hex_warp=0.02, gll_shape=np.array((3, 3, 2)),
         gll_lo=0.04, gll_hi=0.95, gll_warp=0.01, names=np.array(names), on_gll=on_gll, inside=inside, back=back,
         label=np.array(ds.dims[1].label))


def main():
    if len(sys.argv) > 1:  # regenerate selected cases only, e.g. `make_golden_glue.py case_v5`
        for name in sys.argv[1:]:
            globals()[name](SEEDS[name])
        return
    case_v1("o2", (5, 4, 6), 2, 0.03, 500, 7, 101)
    case_v1("o4", (3, 3, 2), 4, 0.02, 250, 23, 102)
    case_v1("2d", (6, 5), 4, 0.03, 400, 5, 103)
    case_v2("o2", (5, 4, 6), 2, 0.03, 700, 11, 201, 1.05)
    case_v2("o4", (3, 2, 3), 4, 0.02, 250, 29, 202, 1.01)
    case_v3(301)
    case_v4(401)
    case_v5(501)
    case_points(601)
    case_gll2gll(701)
    case_layered(801)
    case_query_model(901)
    case_exodus(1001)
    case_gll2gll_2d(1101)
    print("reference calls served by the oracle arithmetic:", refglue.CALLS)


if __name__ == "__main__":
    main()
