"""
Loader that imports the reference's own Python sources (read where they lie under /root/reference,
never copied) with the absent third-party modules replaced by stand-ins, so that the reference's
*own control flow* -- candidate loops V1-V5, fall-backs, de-duplication, gather, fluid fix-up, the
gll_2_gll driver -- can be executed in the build container and its outputs recorded as fixtures.

Used ONLY by tests/golden/make_golden_glue.py (fixture generation).  Nothing here is shipped or
imported by the product, the tests or the bench.

Stand-ins (what is and is not pinned by the fixtures made with them):
  salvus.fem / salvus_fem   closed source, absent.  InverseCoordinateTransformWrapper and
                            GetInterpolationCoefficients are served by the oracle's C restatement
                            (oracle/mm_oracle.c) -> the GLL *arithmetic* stays PARITY UNPINNED; what the
                            fixtures pin is everything the reference does around those two calls.
  pykdtree.kdtree.KDTree    absent.  Exact k-NN served by the oracle's canonical brute force
                            ((d2, index) order; pykdtree's own tie order is unspecified).
  h5py                      absent.  A small in-memory File/Dataset model (paths -> numpy arrays).
  pyexodus                  absent.  A small in-memory `exodus` file model (coords, 1-based connectivity,
                            nodal variables) -- enough for the reference's own io/exodus.py reader.
  multi_mesh*.so            the reference's C library: the reference's own helpers.load_lib() loads the
                            build of its C sources made by oracle/build.py (oracle/_ref), LIB_DIR pointed there.
                            The exodus drivers therefore run WITHOUT any oracle arithmetic: exodus_2_gll is
                            reference Python + reference C end to end.
  names the reference left undefined: interpolator.py:6-7 and io/exodus.py:4-6 comment out the imports of
                            `load_lib`, `Exodus` and the module-level `lib`, and utils.py:200 uses `KDTree`
                            without importing it, so exodus_2_gll / gll_2_exodus / load_exodus /
                            Exodus.get_element_centroid raise NameError as shipped; load_reference_exodus()
                            binds exactly those four names (the first three to the reference's own objects).
  xarray, geographiclib, salvus.mesh, salvus.flow: empty placeholders (never called).
  numpy                     the reference predates numpy 1.24; `np.int` is aliased to int for the
                            duration of the calls.
"""
import importlib
import os
import sys
import types

import numpy as np

REFERENCE = "/root/reference"


# ------------------------------------------------------------------------------------------------
# in-memory h5py
# ------------------------------------------------------------------------------------------------
class _Dim:
    def __init__(self):
        self.label = ""


class FakeDataset:
    def __init__(self, array):
        self.array = np.array(array)
        self.attrs = {}
        self.dims = [_Dim() for _ in range(self.array.ndim)]

    @property
    def shape(self):
        return self.array.shape

    def __getitem__(self, key):
        return np.array(self.array[key])  # h5py hands out copies, never views of the file

    def __setitem__(self, key, value):
        self.array[key] = value

    def __len__(self):
        return len(self.array)


FILES = {}  # filename -> {path: FakeDataset}


class FakeFile:
    def __init__(self, name, mode="r"):
        self.name = str(name)
        if self.name not in FILES:
            if "r" in mode and "+" not in mode:
                raise OSError(f"no such in-memory file {name}")
            FILES[self.name] = {}
        self.store = FILES[self.name]

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def close(self):
        pass

    def __contains__(self, path):
        return path in self.store

    def __getitem__(self, path):
        return self.store[path]

    def __delitem__(self, path):
        del self.store[path]

    def create_dataset(self, name, shape=None, dtype=np.float64, data=None):
        ds = FakeDataset(np.zeros(shape, dtype=dtype) if data is None else data)
        self.store[name] = ds
        return ds

    def keys(self):
        return self.store.keys()


def write_gll_file(name, coordinates, data, params, element_data=None, element_labels=None, global_strings=None):
    """Populate an in-memory Salvus-style GLL file: MODEL/coordinates [E,P,d], MODEL/data [E,F,P]."""
    f = FakeFile(name, "w")
    f.store.clear()
    grp = f.create_dataset("MODEL", data=np.zeros(()))
    for key, val in (global_strings or {}).items():
        grp.attrs[key] = np.bytes_(val)
    f.create_dataset("MODEL/coordinates", data=np.array(coordinates, dtype=np.float64))
    ds = f.create_dataset("MODEL/data", data=np.array(data, dtype=np.float64))
    label = "[ " + " | ".join(params) + " ]"
    ds.attrs["DIMENSION_LABELS"] = ["element", label, "point"]  # str, as load_hdf5_params_to_memory slices it
    if element_data is not None:
        es = f.create_dataset("MODEL/element_data", data=np.array(element_data, dtype=np.float64))
        el = "[ " + " | ".join(element_labels) + " ]"
        es.attrs["DIMENSION_LABELS"] = [b"element", el.encode()]  # bytes, as gll_2_gll decodes it
    return f


# ------------------------------------------------------------------------------------------------
# salvus.fem served by the oracle's C restatement
# ------------------------------------------------------------------------------------------------
def _oracle():
    root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    if root not in sys.path:
        sys.path.insert(0, root)
    from oracle import capi

    return capi


CALLS = {"inverse": 0, "coeffs": 0}


def _inverse(order, dim):
    def f(pnt=None, ctrlNodes=None, **kw):
        capi = _oracle()
        nodes = np.ascontiguousarray(ctrlNodes, dtype=np.float64)
        assert nodes.shape == ((order + 1) ** dim, dim), nodes.shape
        CALLS["inverse"] += 1
        ok, xi = capi.inverse_map_presolved(order, dim, nodes, np.ascontiguousarray(pnt, dtype=np.float64))
        return xi if ok else np.full(dim, np.nan)

    return f


def _coefficients(order, dim):
    def f(ref_coord):
        capi = _oracle()
        CALLS["coeffs"] += 1
        return capi.weights(order, dim, np.ascontiguousarray(ref_coord, dtype=np.float64))

    return f


def _fcts():
    out = []
    for n, P in ((4, 125), (2, 27), (1, 8)):
        out.append((f"__GetInterpolationCoefficients__int_n0_{n}__int_n1_{n}__int_n2_{n}__Matrix_Derive"
                    f"dA_Eigen::Matrix<double, 3, 1>__Matrix_DerivedB_Eigen::Matrix<double, {P}, 1>",
                    _coefficients(n, 3)))
        out.append((f"__InverseCoordinateTransformWrapper__int_n_{n}__int_d_3", _inverse(n, 3)))
    out.append(("__GetInterpolationCoefficients__int_n0_4__int_n1_4__int_n2_0__Matrix_Derive"
                "dA_Eigen::Matrix<double, 2, 1>__Matrix_DerivedB_Eigen::Matrix<double, 25, 1>", _coefficients(4, 2)))
    out.append(("__InverseCoordinateTransformWrapper__int_n_4__int_d_2", _inverse(4, 2)))
    out.append(("__CheckHullWrapper__int_n_4__int_d_3", lambda *a, **k: True))
    return out


class CanonicalKDTree:
    """pykdtree.kdtree.KDTree stand-in: exact k-NN, canonical (d2, index) order."""

    def __init__(self, data, leafsize=16):
        self.data = np.ascontiguousarray(data, dtype=np.float64)

    def query(self, pts, k=1, **kw):
        capi = _oracle()
        pts = np.ascontiguousarray(pts, dtype=np.float64)
        idx, d2 = capi.knn_bruteforce(self.data, pts, k, return_d2=True)
        return np.sqrt(d2), idx.astype(np.int64)


# ------------------------------------------------------------------------------------------------
# in-memory pyexodus
# ------------------------------------------------------------------------------------------------
EXO = {}  # filename -> dict(points [n,3], connectivity0 [E,8] 0-based, nodal {name: values})


def write_exodus_file(name, points, connectivity0, nodal):
    EXO[str(name)] = dict(points=np.array(points, dtype=np.float64),
                          connectivity0=np.array(connectivity0, dtype=np.int64),
                          nodal={k: np.array(v, dtype=np.float64) for k, v in nodal.items()}, names=list(nodal))


class FakeExodusFile:
    def __init__(self, filename, mode="r"):
        self.f = EXO[str(filename)]

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    @property
    def num_dims(self):
        return self.f["points"].shape[1]

    def get_elem_connectivity(self, id=1):
        c = self.f["connectivity0"]
        return c + 1, c.shape[0], c.shape[1]  # exodus is 1-based

    def get_element_variable_names(self):
        return []

    def get_coords(self):
        p = self.f["points"]
        return tuple(p[:, c].copy() for c in range(p.shape[1]))

    def get_node_variable_names(self):
        return list(self.f["names"])

    def get_node_variable_values(self, name, step):
        return self.f["nodal"][name].copy()

    def put_node_variable_name(self, name, index):
        assert self.f["names"][index - 1] == name

    def put_node_variable_values(self, name, step, values):
        self.f["nodal"][name] = np.array(values, dtype=np.float64)


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install_stubs():
    fem = _module("salvus.fem", _fcts=_fcts())
    hyper = _module("salvus.fem.hypercube")

    def wrapper(n=None, d=None, pnt=None, ctrlNodes=None):
        return _inverse(n, d)(pnt=pnt, ctrlNodes=ctrlNodes)

    hyper.InverseCoordinateTransformWrapper = wrapper
    fem.hypercube = hyper
    um = _module("salvus.mesh.unstructured_mesh", UnstructuredMesh=type("UnstructuredMesh", (), {}))
    mesh = _module("salvus.mesh", unstructured_mesh=um)
    flow = _module("salvus.flow")
    _module("salvus", fem=fem, mesh=mesh, flow=flow)
    _module("salvus_fem", _fcts=[(n.replace("n_4__int_d_3", "n_4__int_d_3"), f) for n, f in _fcts()])
    kd = _module("pykdtree.kdtree", KDTree=CanonicalKDTree)
    _module("pykdtree", kdtree=kd)
    _module("h5py", File=FakeFile)
    _module("pyexodus", exodus=FakeExodusFile)
    _module("xarray", Dataset=type("Dataset", (), {}), DataArray=type("DataArray", (), {}))
    gl = _module("geographiclib.geodesic", Geodesic=type("Geodesic", (), {}))
    _module("geographiclib", geodesic=gl)
    if not hasattr(np, "int"):
        np.int = int  # the reference predates numpy 1.24 (interpolator.py:745)


def load_reference():
    """Returns the reference's modules (interpolator, v2_interpolation_tools, utils)."""
    assert os.path.isdir(REFERENCE), "fixture generation needs /root/reference"
    install_stubs()
    for name in [n for n in sys.modules if n == "multi_mesh" or n.startswith("multi_mesh.")]:
        del sys.modules[name]  # make sure the reference's package is the one imported, not the repo's alias
    sys.path.insert(0, REFERENCE)
    try:
        interp = importlib.import_module("multi_mesh.components.interpolator")
        v2 = importlib.import_module("multi_mesh.components.v2_interpolation_tools")
        utils = importlib.import_module("multi_mesh.utils")
    finally:
        sys.path.remove(REFERENCE)
    assert interp.__file__.startswith(REFERENCE), interp.__file__
    return interp, v2, utils


def load_reference_cli():
    """scripts/cli.py (V5 loop); needs click (present) and the same stand-ins."""
    install_stubs()
    sys.path.insert(0, REFERENCE)
    try:
        cli = importlib.import_module("multi_mesh.scripts.cli")
    finally:
        sys.path.remove(REFERENCE)
    assert cli.__file__.startswith(REFERENCE)
    return cli


def load_reference_exodus(interp):
    """Bind the three names the reference left undefined (see the header) to the reference's own objects and
    point its load_lib() at the build of its own C sources."""
    root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    helpers = sys.modules["multi_mesh.helpers"]
    exo = sys.modules["multi_mesh.io.exodus"]
    assert helpers.__file__.startswith(REFERENCE) and exo.__file__.startswith(REFERENCE)
    helpers.LIB_DIR = os.path.join(root, "oracle", "_ref")
    lib = helpers.load_lib()
    exo.lib = lib
    interp.load_lib = helpers.load_lib
    interp.Exodus = exo.Exodus
    sys.modules["multi_mesh.utils"].KDTree = CanonicalKDTree  # utils.py:200 uses KDTree without importing it
    return lib
