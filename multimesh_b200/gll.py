"""
GLL reference-element conventions shared by the host code.

Node index inside an element: a = i + m*j + m*m*k, m = order + 1, i along xi (fastest).
Orders are the ones the reference can dispatch to (interpolator.py:26-57): 1, 2 and 4.
"""
import numpy as np

SQRT_3_7 = float.fromhex("0x1.4f2ec413cb52ap-1")

_NODES = {
    1: (-1.0, 1.0),
    2: (-1.0, 0.0, 1.0),
    4: (-1.0, -SQRT_3_7, 0.0, SQRT_3_7, 1.0),
}

SUPPORTED_ORDERS = tuple(sorted(_NODES))


def gll_nodes(order: int) -> np.ndarray:
    if order not in _NODES:
        raise ValueError(f"GLL order {order} not supported (supported: {SUPPORTED_ORDERS})")
    return np.array(_NODES[order], dtype=np.float64)


def order_from_npoints(n_gll_points: int, dimensions: int) -> int:
    """order = round(P^(1/d)) - 1, as interpolator.py:666-667 / salvus_mesh_reader.py:47-48."""
    order = int(round(n_gll_points ** (1.0 / dimensions))) - 1
    if (order + 1) ** dimensions != n_gll_points:
        raise ValueError(f"{n_gll_points} nodes per element is not (order+1)^{dimensions}")
    return order


def reference_nodes(order: int, dim: int) -> np.ndarray:
    """[P, dim] reference coordinates of the element nodes in storage order."""
    z = gll_nodes(order)
    m = len(z)
    if dim == 2:
        j, i = np.meshgrid(np.arange(m), np.arange(m), indexing="ij")
        return np.stack([z[i.ravel()], z[j.ravel()]], axis=1)
    k, j, i = np.meshgrid(np.arange(m), np.arange(m), np.arange(m), indexing="ij")
    return np.stack([z[i.ravel()], z[j.ravel()], z[k.ravel()]], axis=1)
