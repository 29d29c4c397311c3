"""
Point-cloud generators of the reference's plotter and the interpolations they feed
(multi_mesh/components/plotter.py; line numbers refer to that file).

Only the part of the plotter that sits on the interpolation path is mirrored: building the depth-slice /
cross-section point clouds, pushing them through `interpolate_to_points` (GPU) and the value post-processing
the reference applies before drawing.  The drawing itself (matplotlib / cartopy) is out of scope.
"""
from typing import Tuple

import numpy as np

from ..utils import elliptic_to_geocentric_latitude, greatcircle_points, lat2colat, sph2cart

R_EARTH = 6371000


def _create_depthslice(depth_in_m: float, num: int, lat_extent=(-90.0, 90.0), lon_extent=(-180.0, 180.0)):
    """num x num cloud of [lat, lon, depth_in_m] rows (:159-187)."""
    lat = np.linspace(lat_extent[0], lat_extent[1], num=num)
    lon = np.linspace(lon_extent[0], lon_extent[1], num=num)
    xx, yy = np.meshgrid(lat, lon)
    return np.array((xx.ravel(), yy.ravel(), np.ones_like(yy).ravel() * depth_in_m)).T


def depth_slice_values(mesh, depth_in_km: float, num: int, parameter_to_plot: str, lat_extent=(-90.0, 90.0),
                       lon_extent=(-180.0, 180.0), plot_diff_percentage: bool = False):
    """The [num, num] array plot_depth_slice draws (:86-118): the depth-slice cloud through interpolate_to_points
    (geocentric lat / lon / depth), optionally as percentage deviation from the slice mean."""
    from ..api import interpolate_to_points

    points = _create_depthslice(depth_in_m=depth_in_km * 1000.0, num=num, lat_extent=lat_extent, lon_extent=lon_extent)
    vals = interpolate_to_points(mesh=mesh, points=points, params_to_interp=[parameter_to_plot], make_spherical=False,
                                 geocentric=True).reshape(num, num)
    if plot_diff_percentage:
        lat_mean = np.mean(vals)
        vals = (vals - lat_mean) / lat_mean * 100.0
        if np.max(np.abs(vals)) < 0.1:  # 1-D models (:111-113)
            vals = np.zeros_like(vals)
    return vals


def cross_section_points(point_1_lat, point_1_lng, point_2_lat, point_2_lng, npoints: int, nrads: int,
                         min_depth_in_km: float, max_depth_in_km: float) -> Tuple[np.ndarray, np.ndarray]:
    """xyz cloud [nrads * npoints, 3] of a vertical section along the great circle between two points, and the
    radii (:361-380): geographic latitudes of the great-circle points -> geocentric -> colatitude -> xyz."""
    rads = np.linspace(R_EARTH - max_depth_in_km * 1000, R_EARTH - min_depth_in_km * 1000, nrads)
    lats, lons = greatcircle_points(point_1_lat, point_1_lng, point_2_lat, point_2_lng, npts=npoints).T
    lats = lat2colat(np.array([elliptic_to_geocentric_latitude(v) for v in lats]))
    all_colats, _ = np.meshgrid(lats, rads)
    all_lons, all_rads = np.meshgrid(lons, rads)
    x, y, z = sph2cart(np.deg2rad(all_colats.flatten()), np.deg2rad(all_lons.flatten()), all_rads.ravel())
    return np.array((x, y, z)).T, rads


def cross_section_values(mesh, point_1_lat, point_1_lng, point_2_lat, point_2_lng, param_to_interp: str,
                         npoints: int = 300, nrads: int = 100, min_depth_in_km: float = 0.0,
                         max_depth_in_km: float = 2800.0, relative: bool = True):
    """The [nrads, npoints] array plot_cross_section draws (:381-395): values at the section's points (mesh mapped
    to the sphere), each radius row as percentage deviation from its mean when `relative`."""
    from ..api import interpolate_to_points

    points, _ = cross_section_points(point_1_lat, point_1_lng, point_2_lat, point_2_lng, npoints, nrads,
                                     min_depth_in_km, max_depth_in_km)
    data = interpolate_to_points(mesh, points=points, make_spherical=True, params_to_interp=[param_to_interp])
    data = data.reshape(nrads, npoints)
    if relative:
        for radii in range(nrads):
            data[radii, :] = (data[radii, :] - np.mean(data[radii, :])) / np.mean(data[radii, :]) * 100.0
    return data
