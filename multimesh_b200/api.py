"""
Public API -- drop-in counterpart of multi_mesh/api.py (line numbers refer to that file).

Every entry point keeps the reference's name, positional order and defaults, prints the
reference's "Finished in time: ..." line and forwards to components.interpolator.  Mesh arguments
accept file paths (HDF5 through h5py, or .npz stores) as well as in-memory SalvusMesh / Exodus
objects.  The work runs on the current CUDA device; `threads` arguments are accepted and ignored.
Plotting and `extract_regular_grid` wrappers of the reference are out of scope.
"""
import functools
import pathlib
import time
from typing import List, Union

import numpy as np


def _pick_device(kwargs):
    """`device=` (int / str / torch.device) or `devices=[...]` -- the one additive kwarg of every entry point (SURVEY
    section 5).  One process drives ONE GPU; several GPUs = one process per GPU under torchrun, where every rank calls
    the same entry point and the drivers shard the target points between the ranks (components/interpolator._Source.find)."""
    device = kwargs.pop("device", None)
    devices = kwargs.pop("devices", None)
    if devices is not None:
        devices = list(devices) if isinstance(devices, (list, tuple)) else [devices]
        if len(devices) != 1:
            raise ValueError(
                f"devices={devices}: one process drives one GPU; for {len(devices)} GPUs launch one process per GPU "
                "(`python -m torch.distributed.run --nproc-per-node N script.py`), every rank calling this function "
                "with the same arguments -- the target points are then sharded over the ranks")
        device = devices[0]
    return device


def _on_device(fn):
    """Runs the entry point with `device=` / `devices=[d]` as the current CUDA device."""
    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        import contextlib

        device = _pick_device(kwargs)
        ctx = contextlib.nullcontext()
        if device is not None:
            import torch

            ctx = torch.cuda.device(torch.device(device) if not isinstance(device, int) else device)
        with ctx:
            return fn(*args, **kwargs)

    return wrapper


def _timed(fn):
    fn = _on_device(fn)

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        start = time.time()
        out = fn(*args, **kwargs)
        runtime = time.time() - start
        if runtime >= 60:
            print(f"Finished in time: {runtime / 60} minutes")
        else:
            print(f"Finished in time: {runtime} seconds")
        return out

    return wrapper


@_timed
def query_model(coordinates, model, nelem_to_search=20, parameters="TTI", model_path="MODEL/data",
                coordinates_path="MODEL/coordinates"):
    """Model parameters at [lat, lon, depth_in_m] rows; returns [N, F] (api.py:13-58)."""
    from .components.interpolator import query_model as _impl

    return _impl(coordinates=coordinates, model=model, nelem_to_search=nelem_to_search,
                 model_path=model_path, coordinates_path=coordinates_path)


@_timed
def exodus_2_gll(mesh, gll_model, gll_order=4, dimensions=3, nelem_to_search=20, parameters="TTI",
                 model_path="MODEL/data", coordinates_path="MODEL/coordinates"):
    """Exodus (nodal HEX8) model -> GLL model, 3-D only (api.py:61-103)."""
    from .components.interpolator import exodus_2_gll as _impl

    _impl(mesh, gll_model, gll_order, dimensions, nelem_to_search, parameters, model_path, coordinates_path)


@_timed
def gll_2_gll(from_gll, to_gll, nelem_to_search=20, parameters="TTI", from_model_path="MODEL/data",
              to_model_path="MODEL/data", from_coordinates_path="MODEL/coordinates",
              to_coordinates_path="MODEL/coordinates", gradient=False, stored_array=None):
    """GLL model -> GLL model; the target file is updated in place (api.py:106-155)."""
    from .components.interpolator import gll_2_gll as _impl

    _impl(from_gll, to_gll, nelem_to_search=nelem_to_search, parameters=parameters,
          from_model_path=from_model_path, to_model_path=to_model_path,
          from_coordinates_path=from_coordinates_path, to_coordinates_path=to_coordinates_path,
          gradient=gradient, stored_array=stored_array)


@_timed
def gll_2_gll_layered(from_gll: Union[str, pathlib.Path], to_gll: Union[str, pathlib.Path],
                      layers: Union[str, List[int]], nelem_to_search: int = 20,
                      parameters: Union[str, List[str]] = "ISO", stored_array: Union[str, pathlib.Path] = None,
                      make_spherical: bool = False):
    """Layer-restricted GLL -> GLL interpolation (api.py:158-215)."""
    from .components.interpolator import gll_2_gll_layered as _impl

    _impl(from_gll=from_gll, to_gll=to_gll, layers=layers, nelem_to_search=nelem_to_search,
          parameters=parameters, stored_array=stored_array, make_spherical=make_spherical)


@_timed
def gll_2_gll_layered_multi(from_gll: Union[str, pathlib.Path], to_gll: Union[str, pathlib.Path],
                            layers: Union[List[int], str] = "nocore", nelem_to_search: int = 20,
                            parameters: Union[List[str], str] = "all", threads: int = None,
                            stored_array: Union[str, pathlib.Path] = None, make_spherical: bool = False):
    """Layered interpolation, layers processed back to back on the GPU (api.py:218-274)."""
    from .components.interpolator import gll_2_gll_layered_multi as _impl

    _impl(from_gll=from_gll, to_gll=to_gll, layers=layers, nelem_to_search=nelem_to_search,
          parameters=parameters, threads=threads, stored_array=stored_array, make_spherical=make_spherical)


@_timed
def gll_2_exodus(gll_model, exodus_model, gll_order=4, dimensions=3, nelem_to_search=20, parameters="TTI",
                 model_path="MODEL/data", coordinates_path="MODEL/coordinates", gradient=False):
    """GLL model -> Exodus nodal fields (api.py:277-317)."""
    from .components.interpolator import gll_2_exodus as _impl

    _impl(gll_model, exodus_model, gll_order, dimensions, nelem_to_search, parameters, model_path,
          coordinates_path, gradient)


@_on_device
def interpolate_to_points(mesh, points, params_to_interp, make_spherical=False, geocentric=False):
    """Mesh -> point cloud (xyz, or lat/lon/depth when `geocentric`); returns [N, F]
    (api.py:320-350)."""
    if geocentric:
        from .utils import latlondepth_to_xyz

        points = latlondepth_to_xyz(points)
    from .components.interpolator import interpolate_to_points as _impl

    return _impl(mesh=mesh, points=points, params_to_interp=params_to_interp, make_spherical=make_spherical)


@_on_device
def interpolate_to_mesh(old_mesh, new_mesh, params_to_interp=["VSV", "VSH", "VPV", "VPH"]):
    """Map both meshes to the sphere, interpolate old -> new at the new mesh's nodes, restore the
    coordinates; points that are not found get zero (api.py:353-393).  Meshes are SalvusMesh
    objects or paths (the reference needs salvus' UnstructuredMesh, which is not available)."""
    from .components.interpolator import _as_salvus_mesh, interpolate_to_points as _impl, map_to_sphere

    old_mesh = _as_salvus_mesh(old_mesh)
    new_mesh = _as_salvus_mesh(new_mesh)
    orig_old, orig_new = np.copy(old_mesh.points), np.copy(new_mesh.points)
    map_to_sphere(old_mesh)
    map_to_sphere(new_mesh)
    vals = _impl(old_mesh, new_mesh.points.reshape(-1, new_mesh.points.shape[-1]), params_to_interp)
    for i, param in enumerate(params_to_interp):
        new_mesh.attach_field(param, vals[:, i].reshape(new_mesh.nelem, new_mesh.n_gll_points))
    old_mesh.points[...] = orig_old
    new_mesh.points[...] = orig_new
    return new_mesh


@_timed
def gll_2_gll_layered_multi_two(from_gll: Union[str, pathlib.Path], to_gll: Union[str, pathlib.Path],
                                layers: Union[List[int], str], nelem_to_search: int = 30,
                                parameters: Union[List[str], str] = "all",
                                stored_array: Union[str, pathlib.Path] = None, make_spherical: bool = False,
                                tolerance: float = 1.05):
    """Layered interpolation with tolerance + snap-to-nearest location (api.py:645-699)."""
    from .components.interpolator import gll_2_gll_layered_multi_two as _impl

    _impl(from_gll=from_gll, to_gll=to_gll, layers=layers, nelem_to_search=nelem_to_search,
          parameters=parameters, stored_array=stored_array, make_spherical=make_spherical, tolerance=tolerance)
