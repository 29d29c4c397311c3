"""
Tier-A oracle: a slow, readable, *independently written* numpy restatement of the hot path,
shaped like the reference's Python (per-point loops, explicit coefficient arrays, numpy sums).
TEST INFRASTRUCTURE ONLY.  Used on small cases to cross-check oracle/mm_oracle.c, which uses a
different (sum-factorised) evaluation order; agreement is to roundoff, not bit-for-bit.

PARITY UNPINNED for the GLL arithmetic: salvus.fem is closed source and absent (see
oracle/mm_oracle.c header).  The functions below restate the published mathematics at the
reference's call sites:
    inverse_transform   interpolator.py:1370-1386
    get_coefficients    interpolator.py:1337-1347
    boundary_box_check  interpolator.py:1350-1367
    _check_if_inside_element (V1)            interpolator.py:1409-1473
    get_element_weights.check_inside (V2)    interpolator.py:1181-1233
    get_element_weights_layered (V3)         interpolator.py:1271-1297
"""
import numpy as np
from numpy.polynomial import legendre as _leg


def gll_nodes(order):
    """Gauss-Lobatto-Legendre nodes: +-1 and the roots of P_n'(x)."""
    if order == 1:
        return np.array([-1.0, 1.0])
    c = np.zeros(order + 1)
    c[order] = 1.0
    inner = _leg.legroots(_leg.legder(c))
    return np.concatenate([[-1.0], np.sort(inner), [1.0]])


def lagrange_1d(nodes, x):
    """L_i(x) and L_i'(x) straight from the product formula."""
    m = len(nodes)
    L = np.ones(m)
    dL = np.zeros(m)
    for i in range(m):
        others = [j for j in range(m) if j != i]
        den = np.prod([nodes[i] - nodes[j] for j in others])
        L[i] = np.prod([x - nodes[j] for j in others]) / den
        s = 0.0
        for q in others:
            s += np.prod([x - nodes[j] for j in others if j != q])
        dL[i] = s / den
    return L, dL


def coefficients(order, xi):
    """Tensor-product weights, node a = i + m*j + m*m*k (i along xi fastest)."""
    z = gll_nodes(order)
    Ls = [lagrange_1d(z, x)[0] for x in xi]
    if len(xi) == 2:
        return np.einsum("j,i->ji", Ls[1], Ls[0]).ravel()
    return np.einsum("k,j,i->kji", Ls[2], Ls[1], Ls[0]).ravel()


def _weights_and_grads(order, xi):
    z = gll_nodes(order)
    LD = [lagrange_1d(z, x) for x in xi]
    d = len(xi)
    if d == 2:
        w = np.einsum("j,i->ji", LD[1][0], LD[0][0]).ravel()
        g0 = np.einsum("j,i->ji", LD[1][0], LD[0][1]).ravel()
        g1 = np.einsum("j,i->ji", LD[1][1], LD[0][0]).ravel()
        return w, np.stack([g0, g1], axis=1)
    w = np.einsum("k,j,i->kji", LD[2][0], LD[1][0], LD[0][0]).ravel()
    g0 = np.einsum("k,j,i->kji", LD[2][0], LD[1][0], LD[0][1]).ravel()
    g1 = np.einsum("k,j,i->kji", LD[2][0], LD[1][1], LD[0][0]).ravel()
    g2 = np.einsum("k,j,i->kji", LD[2][1], LD[1][0], LD[0][0]).ravel()
    return w, np.stack([g0, g1, g2], axis=1)


def forward_map(order, nodes, xi):
    return coefficients(order, xi) @ nodes


def inverse_transform(point, gll_points, order, maxit=50, tol=1e-13):
    """Newton from xi = 0 on the order-n map; returns NaNs when it does not converge."""
    d = gll_points.shape[1]
    xi = np.zeros(d)
    for _ in range(maxit):
        w, g = _weights_and_grads(order, xi)
        x = w @ gll_points
        J = gll_points.T @ g  # J[c, s] = dx_c / dxi_s
        try:
            delta = np.linalg.solve(J, point - x)
        except np.linalg.LinAlgError:
            return np.full(d, np.nan)
        if not np.all(np.abs(delta) <= 1e10):
            return np.full(d, np.nan)
        xi = xi + delta
        if np.max(np.abs(delta)) <= tol:
            return xi
    return np.full(d, np.nan)


def boundary_box_check(point, gll_points):
    p_min, p_max = gll_points.min(axis=0), gll_points.max(axis=0)
    if (point >= p_min).all() and (point <= p_max).all():
        return True, 0.0
    center = np.mean(gll_points, axis=0)
    return False, float(np.linalg.norm(point - center))


MAGIC_XI = np.array([0.645, -0.5, 0.22])


def check_if_inside_element_v1(gll_model, nearest_elements, point, order, ignore_hard_elements=True):
    d = gll_model.shape[2]
    dist = np.zeros(len(nearest_elements))
    inside = np.zeros(len(nearest_elements), dtype=bool)
    for _i, element in enumerate(nearest_elements):
        gp = gll_model[element]
        inside[_i], dist[_i] = boundary_box_check(point, gp)
        if inside[_i]:
            ref = inverse_transform(point, gp, order)
            if np.any(np.isnan(ref)):
                continue
            if np.all(np.abs(ref) <= 1.04):
                return element, ref
    if np.any(inside):
        ind = np.where(dist == np.min(dist[np.where(inside)]))[0][0]
    else:
        ind = np.where(dist == np.min(dist))[0][0]
    element = nearest_elements[ind]
    ref = inverse_transform(point, gll_model[element], order)
    if np.any(np.isnan(ref)):
        if not ignore_hard_elements:
            raise ValueError("Can't find an appropriate element.")
        ref = MAGIC_XI[:d].copy()
    if np.any(np.abs(ref) >= 1.04):
        ref = MAGIC_XI[:d].copy()
    return element, ref


def check_inside_v2(gll_points, nearest_elements, point, order, tolerance=1.05, snap_to_nearest=False):
    d = gll_points.shape[2]
    best = 10e9
    best_elem = 0
    for element in nearest_elements:
        ref = inverse_transform(point, gll_points[element], order)
        if np.any(np.isnan(ref)):
            continue
        if np.max(np.abs(ref)) < np.max(np.abs(best)):
            best = ref
            best_elem = element
        if np.all(np.abs(ref) < tolerance):
            return element, ref
    if snap_to_nearest:
        ref = np.clip(best, -1.02, 1.02) * np.ones(d)
        return best_elem, ref
    return -1, None


def check_inside_v3(gll_points, nearest_elements, point, order):
    for element in nearest_elements:
        ref = inverse_transform(point, gll_points[element], order)
        if np.any(np.isnan(ref)):
            continue
        if np.all(np.abs(ref) < 1.03):
            return element, ref
    return -1, None


def knn(data, pts, k):
    """Brute-force k-NN in the canonical (d2, index) order via lexsort."""
    out = np.zeros((len(pts), k), dtype=np.int64)
    for n, p in enumerate(pts):
        diff = p - data
        d2 = diff[:, 0] * diff[:, 0] + diff[:, 1] * diff[:, 1]
        if data.shape[1] == 3:
            d2 = d2 + diff[:, 2] * diff[:, 2]
        out[n] = np.lexsort((np.arange(len(data)), d2))[:k]
    return out


def gather(fields, elements, coeffs):
    """values[n, f] = sum_a fields[elem_n, f, a] * coeffs[n, a]  (interpolator.py:814-826)."""
    return np.sum(fields[elements] * coeffs[:, None, :], axis=2)


def unique_points(points):
    """utils.get_unique_points of the reference (utils.py:484-492): np.unique over the flattened GLL nodes."""
    allp = points.reshape(points.shape[0] * points.shape[1], points.shape[2])
    u, inv = np.unique(allp, return_inverse=True, axis=0)
    return u, inv.reshape(-1)
