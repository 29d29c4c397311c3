"""
Host-side helpers on the interpolation path; counterparts of the functions of the same name in
multi_mesh/utils.py (line numbers below refer to that file).  Visualisation / xarray / geodesy
helpers of the reference are out of scope.
"""
from typing import List, Union

import numpy as np

from .io.store import open_store

R_EARTH = 6371000.0

_PRESETS = {
    "TTI": ["VPV", "VPH", "VSV", "VSH", "RHO", "ETA", "QKAPPA", "QMU"],  # :172-182
    "ISO": ["QKAPPA", "QMU", "RHO", "VP", "VS"],  # :183-184
}


def pick_parameters(parameters):
    """'TTI' / 'ISO' presets, anything else is passed through (:171-188)."""
    if isinstance(parameters, str) and parameters in _PRESETS:
        return list(_PRESETS[parameters])
    return parameters


def load_hdf5_params_to_memory(gll: str, model: str, coordinates: str):
    """(points [E,P,d] f64, data [E,F,P], params) with 'grad' stripped from the labels (:206-217)."""
    with open_store(gll, "r") as st:
        points = np.array(st.read(coordinates), dtype=np.float64)
        data = np.array(st.read(model))
        params = [p.replace("grad", "") for p in st.labels(model)]
    return points, data, params


def load_hdf5_params_to_device(gll: str, model: str, coordinates: str, device=None):
    """Device twin of load_hdf5_params_to_memory: (points [E,P,d], data [E,F,P]) as CUDA tensors + params.  Each
    array is read from the file straight into pinned memory and copied asynchronously (io/staging.py), the
    coordinates first, so that geometry and index can be built while the field array is still in flight.
    Returns (staged, params): `staged[coordinates]` / `staged[model]` wait (stream-ordered) for their copy."""
    from .io.staging import stage_arrays
    from .kdtree import _device

    with open_store(gll, "r") as st:
        params = [p.replace("grad", "") for p in st.labels(model)]
        staged = stage_arrays(st, [coordinates, model], _device(device))
    return staged, params


def remove_and_create_empty_dataset(gll_model, parameters: list, model: str, coordinates: str):
    """Replace `model` by an empty [E, len(parameters), P] float64 dataset with fresh dimension
    labels (:137-168).  `gll_model` is an open store."""
    E, P = gll_model.shape(coordinates)[:2]
    gll_model.write(model, np.zeros((E, len(parameters), P), dtype=np.float64))
    gll_model.set_labels(model, list(parameters))


def _layer_field(mesh):
    fields = mesh.get_elemental_fields() if hasattr(mesh, "get_elemental_fields") else mesh.elemental_fields
    return fields


def _assess_layers(mesh, layers: Union[List[int], str, int]):
    """Resolve a layer request into the list of numerical layers (descending, so that moho_idx
    indexes from the surface) and whether masking is needed (:382-440)."""
    fields = _layer_field(mesh)
    mesh_layers = np.sort(np.unique(fields["layer"]))[::-1].astype(int)
    if isinstance(layers, (list, tuple, np.ndarray)):
        if np.max(layers) > np.max(mesh_layers) or np.min(layers) < np.min(mesh_layers):
            raise ValueError("Requested layers not in mesh")
        return list(layers), set(mesh_layers.tolist()) != set(int(x) for x in layers)
    if isinstance(layers, (int, np.integer)):
        if layers not in mesh_layers:
            raise ValueError("Requested layer not in mesh")
        return [int(layers)], True
    available = ["all", "crust", "mantle", "core", "nocore"]
    if not isinstance(layers, str):
        raise ValueError(f"Input for layers needs to be a list of one of: {available}")
    if layers == "all":
        return mesh_layers, False
    if layers not in available:
        raise ValueError(f"Only allowed string layer inputs are: {available}")
    if layers == "crust":
        return mesh_layers[: int(mesh.global_strings["moho_idx"])], True
    fluid = np.where(fields["fluid"] == 1)[0]
    if fluid.size == 0:
        # no fluid element: the reference would raise IndexError; treat as "no core present"
        o_core_idx = len(mesh_layers)
    else:
        o_core_idx = int(np.where(mesh_layers == fields["layer"][fluid[0]])[0][0])
    if layers == "mantle":
        return mesh_layers[int(mesh.global_strings["moho_idx"]):o_core_idx], True
    if layers == "core":
        return mesh_layers[o_core_idx:], True
    return mesh_layers[:o_core_idx], True  # nocore


def _create_mask(mesh, layers):
    """{str(layer): bool[E]} (:355-379)."""
    lay = _layer_field(mesh)["layer"]
    return {str(layer): (lay == layer) for layer in layers}, layers


def create_layer_mask(mesh, layers):
    layers, _ = _assess_layers(mesh=mesh, layers=layers)
    return _create_mask(mesh=mesh, layers=layers)


def _unique_rows(allp, device=None, as_numpy=True):
    """K4 on the device (ops.unique_points = mm_unique_points): the rows of np.unique(allp, axis=0) in the same
    lexicographic order, and the inverse map.  There is no host fallback."""
    import torch

    from . import ops
    from .kdtree import _device

    dev = _device(device)
    t = allp if isinstance(allp, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(allp, dtype=np.float64))
    u, inv = ops.unique_points(t.to(dev, dtype=torch.float64).contiguous())
    if as_numpy:
        return u.cpu().numpy(), inv.cpu().numpy().astype(np.int64)
    return u, inv


def get_unique_points(points, mesh=False, layers=None, device=None, as_numpy=True):
    """Lexicographic unique rows + inverse (np.unique(axis=0, return_inverse=True)), for a
    coordinate array [E,P,d] or, per layer, for a mesh object (:465-515).  Computed on the GPU (K4);
    `as_numpy=False` leaves (unique, inverse int32) on the device for the drivers."""
    if not hasattr(points, "get_element_nodes"):
        allp = points.reshape(points.shape[0] * points.shape[1], points.shape[2])
        return _unique_rows(allp, device, as_numpy)
    layers, _ = _assess_layers(mesh=points, layers=layers)
    mask, _ = _create_mask(mesh=points, layers=layers)
    unique_points = {}
    for layer in layers:
        nodes = points.get_element_nodes()[mask[str(layer)]]
        unique_points[str(layer)] = _unique_rows(nodes.reshape(nodes.shape[0] * nodes.shape[1], nodes.shape[2]),
                                                 device, as_numpy)
    return unique_points, mask, layers


def lat2colat(lat):
    return 90.0 - lat


def latlondepth_to_xyz(latlondepth: np.ndarray):
    """[lat, lon, depth_m] rows -> geocentric xyz on a 6371 km sphere (:526-542)."""
    r = R_EARTH - latlondepth[:, 2]
    colat = np.deg2rad(lat2colat(latlondepth[:, 0]))
    lon = np.deg2rad(latlondepth[:, 1])
    return np.array([r * np.sin(colat) * np.cos(lon), r * np.sin(colat) * np.sin(lon),
                     r * np.cos(colat)]).T


def load_exodus(file, find_centroids=True):
    """Exodus object (+ KD-tree over its element centroids) (:191-203)."""
    from .io.exodus import Exodus
    from .kdtree import KDTree

    exodus = file if isinstance(file, Exodus) else Exodus(file)
    if not find_centroids:
        return exodus
    return exodus, KDTree(exodus.get_element_centroid())


# ----------------------------------------------------------------------------------------------
# geodesy helpers used by the point-cloud generators of the plotter (utils.py:95-134, 545-604)
# ----------------------------------------------------------------------------------------------
def sph2cart(col, lon, rad):
    """(colatitude [rad], longitude [rad], radius) -> x, y, z (utils.py:577-594)."""
    col, lon, rad = np.asarray(col), np.asarray(lon), np.asarray(rad)
    if (0 > col).any() or (col > np.pi).any():
        raise ValueError("Colatitude must be in range [0, pi].")
    return rad * np.sin(col) * np.cos(lon), rad * np.sin(col) * np.sin(lon), rad * np.cos(col)


def cart2sph(x, y, z):
    """x, y, z -> (colatitude, longitude, radius); the centre maps to colatitude pi/2 (utils.py:597-617)."""
    x, y, z = np.asarray(x), np.asarray(y), np.asarray(z)
    r = np.sqrt(x ** 2 + y ** 2 + z ** 2)
    with np.errstate(invalid="ignore", divide="ignore"):
        c = np.nan_to_num(np.divide(z, r))
    return np.arccos(c), np.arctan2(y, x), r


def elliptic_to_geocentric_latitude(lat, axis_a=6378137.0, axis_b=6356752.314245):
    """WGS84 geographic -> geocentric latitude in degrees: atan((1 - e^2) tan(lat)).  The reference imports this
    from LASIF (components/plotter.py:8), which is not installed here; this is LASIF's published formula."""
    f = (axis_a - axis_b) / axis_a
    e2 = 2.0 * f - f ** 2
    if abs(lat) < 1e-6 or abs(lat - 90.0) < 1e-6 or abs(lat + 90.0) < 1e-6:
        return lat
    return float(np.rad2deg(np.arctan((1.0 - e2) * np.tan(np.deg2rad(lat)))))


def greatcircle_points(point_1_lat, point_1_lng, point_2_lat, point_2_lng, npts=101):
    """npts points [lat, lon] from point 1 towards point 2, the i-th at i / npts of the way (the last point stops
    one step short of point 2, as in the reference, utils.py:545-574).  The reference walks the WGS84 geodesic with
    geographiclib, which is not installed here; this walks the great circle of the sphere (differences are below
    0.2 degrees and only move the plotted section slightly)."""
    if npts < 3:
        raise Exception("You should supply at least 3 points")
    a = np.deg2rad([point_1_lat, point_1_lng])
    b = np.deg2rad([point_2_lat, point_2_lng])
    va = np.array([np.cos(a[0]) * np.cos(a[1]), np.cos(a[0]) * np.sin(a[1]), np.sin(a[0])])
    vb = np.array([np.cos(b[0]) * np.cos(b[1]), np.cos(b[0]) * np.sin(b[1]), np.sin(b[0])])
    omega = np.arccos(np.clip(va @ vb, -1.0, 1.0))
    out = []
    for i in range(npts):
        t = i / float(npts)
        v = va if omega < 1e-15 else (np.sin((1 - t) * omega) * va + np.sin(t * omega) * vb) / np.sin(omega)
        out.append([np.rad2deg(np.arcsin(np.clip(v[2], -1.0, 1.0))), np.rad2deg(np.arctan2(v[1], v[0]))])
    return np.array(out)
