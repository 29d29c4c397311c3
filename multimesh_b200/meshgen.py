"""
Synthetic mesh generators (host, numpy) in the reference's data layout:

    MODEL/coordinates  [E, P, d]  float64      (salvus_mesh_reader.py:38-39)
    MODEL/data         [E, F, P]  float64      (salvus_mesh_reader.py:90-97)
    element_data       layer / fluid per element (utils.py:372-379,427-429)

Used by the tests, the smoke test and bench.py (inputs S1-S5).  There are no
Salvus/Exodus files to read in this environment, so these stand in for them.
"""
import numpy as np

from .gll import gll_nodes

R_EARTH = 6371000.0


def _elem_origin_grid(shape):
    """element index e = ex + nx*(ey + ny*ez)  ->  integer element coordinates [E, d]."""
    d = len(shape)
    if d == 2:
        ey, ex = np.meshgrid(np.arange(shape[1]), np.arange(shape[0]), indexing="ij")
        return np.stack([ex.ravel(), ey.ravel()], axis=1)
    ez, ey, ex = np.meshgrid(np.arange(shape[2]), np.arange(shape[1]), np.arange(shape[0]),
                             indexing="ij")
    return np.stack([ex.ravel(), ey.ravel(), ez.ravel()], axis=1)


def box_mesh(shape, order, lo=None, hi=None, warp=0.0, dtype=np.float64):
    """Structured quad/hex GLL mesh on the box [lo, hi].

    shape : elements per axis, length d (2 or 3)
    warp  : amplitude (fraction of the box) of a smooth sinusoidal deformation applied to every
            GLL node, which makes the order-n geometry genuinely curved (non-trilinear).
    Returns coordinates [E, P, d].
    """
    shape = tuple(int(s) for s in shape)
    d = len(shape)
    lo = np.zeros(d) if lo is None else np.asarray(lo, dtype=np.float64)
    hi = np.ones(d) if hi is None else np.asarray(hi, dtype=np.float64)
    z = gll_nodes(order)
    m = len(z)
    t = 0.5 * (z + 1.0)  # node offsets inside an element, in [0, 1]
    eorg = _elem_origin_grid(shape)  # [E, d]
    E = eorg.shape[0]
    P = m ** d
    coords = np.empty((E, P, d), dtype=dtype)
    # local node a = i + m*j + m*m*k
    if d == 2:
        jj, ii = np.meshgrid(np.arange(m), np.arange(m), indexing="ij")
        loc = [ii.ravel(), jj.ravel()]
    else:
        kk, jj, ii = np.meshgrid(np.arange(m), np.arange(m), np.arange(m), indexing="ij")
        loc = [ii.ravel(), jj.ravel(), kk.ravel()]
    for c in range(d):
        u = (eorg[:, c:c + 1] + t[loc[c]][None, :]) / shape[c]  # in [0, 1]
        coords[:, :, c] = u
    if warp:
        u = coords.copy()
        for c in range(d):
            o = (c + 1) % d
            coords[:, :, c] = u[:, :, c] + warp * np.sin(2 * np.pi * u[:, :, o]) * np.sin(
                np.pi * u[:, :, c])
    for c in range(d):
        coords[:, :, c] = lo[c] + (hi[c] - lo[c]) * coords[:, :, c]
    return coords


def analytic_fields(coords, names, scale=None):
    """Smooth analytic nodal fields in MODEL/data layout [E, F, P]."""
    E, P, d = coords.shape
    lo = coords.reshape(-1, d).min(axis=0)
    hi = coords.reshape(-1, d).max(axis=0)
    span = np.where(hi > lo, hi - lo, 1.0) if scale is None else scale
    u = (coords - lo) / span
    x, y = u[..., 0], u[..., 1]
    zc = u[..., 2] if d == 3 else 0.0 * x
    out = np.empty((E, len(names), P))
    for f, name in enumerate(names):
        key = name.upper().replace("GRAD", "")
        if key in ("VP", "VPV", "VPH"):
            v = 5000.0 + 800.0 * np.sin(2 * np.pi * x) * np.cos(2 * np.pi * y) + 300.0 * zc
        elif key in ("VS", "VSV", "VSH"):
            v = (5000.0 + 800.0 * np.sin(2 * np.pi * x) * np.cos(2 * np.pi * y) + 300.0 * zc) / np.sqrt(3.0)
        elif key == "RHO":
            v = 2600.0 + 300.0 * x * y + 150.0 * zc * zc
        elif key == "QKAPPA":
            v = 57823.0 + 100.0 * np.cos(np.pi * (x + y + zc))
        elif key == "QMU":
            v = 600.0 - 80.0 * x + 40.0 * y * zc
        elif key == "ETA":
            v = 1.0 + 0.05 * np.sin(np.pi * x) * np.sin(np.pi * y)
        else:
            v = 1.0 + f + x + 2.0 * y + 3.0 * zc
        out[:, f, :] = v
    return out


def polynomial_field(coords, degree, rng):
    """A random polynomial with per-axis degree <= `degree`; reproduced exactly by order >= degree
    interpolation on affine elements (test property 3)."""
    E, P, d = coords.shape
    coef = rng.uniform(-1.0, 1.0, size=(degree + 1,) * d)
    x = [coords[..., c] for c in range(d)]
    v = np.zeros((E, P))
    for idx in np.ndindex(*coef.shape):
        term = coef[idx]
        for c in range(d):
            term = term * x[c] ** idx[c]
        v += term
    return v, coef


# ----------------------------------------------------------------------------------------------
# HEX8 nodal ("exodus-style") mesh: points [Np, 3] + connectivity [E, 8], 0-based,
# Exodus vertex order (bottom face counter-clockwise, then top face), io/exodus.py:36-46.
# ----------------------------------------------------------------------------------------------
def hex8_mesh(shape, lo=None, hi=None, warp=0.0):
    nx, ny, nz = (int(s) for s in shape)
    lo = np.zeros(3) if lo is None else np.asarray(lo, dtype=np.float64)
    hi = np.ones(3) if hi is None else np.asarray(hi, dtype=np.float64)
    gz, gy, gx = np.meshgrid(np.arange(nz + 1), np.arange(ny + 1), np.arange(nx + 1), indexing="ij")
    u = np.stack([gx.ravel() / nx, gy.ravel() / ny, gz.ravel() / nz], axis=1)
    if warp:
        v = u.copy()
        for c in range(3):
            o = (c + 1) % 3
            u[:, c] = v[:, c] + warp * np.sin(2 * np.pi * v[:, o]) * np.sin(np.pi * v[:, c])
    points = lo + (hi - lo) * u

    def nid(ix, iy, iz):
        return ix + (nx + 1) * (iy + (ny + 1) * iz)

    ez, ey, ex = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    ex, ey, ez = ex.ravel(), ey.ravel(), ez.ravel()
    conn = np.stack([
        nid(ex, ey, ez), nid(ex + 1, ey, ez), nid(ex + 1, ey + 1, ez), nid(ex, ey + 1, ez),
        nid(ex, ey, ez + 1), nid(ex + 1, ey, ez + 1), nid(ex + 1, ey + 1, ez + 1),
        nid(ex, ey + 1, ez + 1),
    ], axis=1).astype(np.int64)
    return np.ascontiguousarray(points), np.ascontiguousarray(conn)


# ----------------------------------------------------------------------------------------------
# Cubed-sphere spherical shell with radial layers (input S3)
# ----------------------------------------------------------------------------------------------
_FACES = (
    # (axis of the normal, sign, axes of the two tangent directions)
    (0, +1.0, 1, 2), (0, -1.0, 2, 1), (1, +1.0, 2, 0), (1, -1.0, 0, 2), (2, +1.0, 0, 1), (2, -1.0, 1, 0),
)


def shell_mesh(n_lat, layers, order, r_earth=R_EARTH):
    """Cubed-sphere shell: 6 chunks x (n_lat x n_lat) lateral cells x radial elements.

    layers : list of (r_bottom, r_top, n_radial_elements, layer_id, fluid_flag); radii in metres.
             Stacked bottom-up; the thin-crust / thick-mantle contrast of global meshes is
             expressed through the radii.
    Every GLL node is placed on its exact sphere (equiangular gnomonic projection), so the
    order-n geometry is curved.  Returns (coords [E,P,3], elemental {layer, fluid}, z_node_1D [E,P]).
    Element order: radial index fastest within a lateral cell?  No: e = cell + ncell * irad_global,
    i.e. whole shells are contiguous, bottom shell first.
    """
    z = gll_nodes(order)
    m = len(z)
    t = 0.5 * (z + 1.0)
    kk, jj, ii = np.meshgrid(np.arange(m), np.arange(m), np.arange(m), indexing="ij")
    li, lj, lk = ii.ravel(), jj.ravel(), kk.ravel()
    # lateral cells of all six faces
    cv, cu = np.meshgrid(np.arange(n_lat), np.arange(n_lat), indexing="ij")
    cu, cv = cu.ravel(), cv.ravel()
    dirs = []
    for (ax, sgn, a1, a2) in _FACES:
        u = (cu[:, None] + t[li][None, :]) / n_lat * 2.0 - 1.0  # [-1, 1]
        v = (cv[:, None] + t[lj][None, :]) / n_lat * 2.0 - 1.0
        a = np.tan(u * (np.pi / 4.0))
        b = np.tan(v * (np.pi / 4.0))
        vec = np.empty(a.shape + (3,))
        vec[..., ax] = sgn
        vec[..., a1] = a
        vec[..., a2] = b * sgn
        vec /= np.linalg.norm(vec, axis=-1, keepdims=True)
        dirs.append(vec)
    dirs = np.concatenate(dirs, axis=0)  # [ncell, P, 3]
    ncell = dirs.shape[0]
    coords_l, layer_l, fluid_l, r1d_l = [], [], [], []
    for (r0, r1, nrad, lid, fluid) in layers:
        for ir in range(nrad):
            rb = r0 + (r1 - r0) * ir / nrad
            rt = r0 + (r1 - r0) * (ir + 1) / nrad
            r = rb + (rt - rb) * t[lk]  # [P]
            coords_l.append(dirs * r[None, :, None])
            r1d_l.append(np.broadcast_to(r[None, :] / r_earth, (ncell, m ** 3)).copy())
            layer_l.append(np.full(ncell, lid, dtype=np.float64))
            fluid_l.append(np.full(ncell, float(fluid)))
    coords = np.ascontiguousarray(np.concatenate(coords_l, axis=0))
    elemental = {"layer": np.concatenate(layer_l), "fluid": np.concatenate(fluid_l)}
    z1d = np.concatenate(r1d_l, axis=0)
    return coords, elemental, z1d


def default_shell_layers(scale=1):
    """Three radial layers (mantle, lower crust, thin upper crust); none fluid (the `nocore`
    configuration of gll_2_gll_layered_multi, api.py:218-227)."""
    return [
        (3480e3, 6291e3, 6 * scale, 3, 0),
        (6291e3, 6346e3, 1 * scale, 2, 0),
        (6346e3, 6371e3, 1 * scale, 1, 0),
    ]


def random_points_in_box(n, lo, hi, seed=1234):
    rng = np.random.default_rng(seed)
    lo = np.asarray(lo, dtype=np.float64)
    hi = np.asarray(hi, dtype=np.float64)
    return lo + (hi - lo) * rng.random((n, len(lo)))
