"""
Library loading helper -- the counterpart of multi_mesh/helpers.py:29-84.

`load_lib()` returns the cached ctypes handle of `lib/multi_mesh*.so`; `lib.centroid` and
`lib.triLinearInterpolator` keep the reference's argument meaning (numpy arrays, C-contiguous)
but are declared with the `long long` types the C signatures really use (the reference declares
`c_int`).  Raises ValueError when the shared library is missing.
"""
import ctypes as C

import numpy as np

from . import _lib

LIB_DIR = _lib.LIB_DIR
cache = []


class _NumpyLib:
    """Thin adaptor so the two legacy entry points accept numpy arrays like the reference's
    ndpointer-typed functions do."""

    def __init__(self, lib):
        self._lib = lib

    @staticmethod
    def _p(a, dtype):
        if not (isinstance(a, np.ndarray) and a.dtype == dtype and a.flags["C_CONTIGUOUS"]):
            raise TypeError(f"expected a C-contiguous numpy array of dtype {np.dtype(dtype)}")
        return a.ctypes.data_as(C.c_void_p)

    def centroid(self, ndim, nelem, npointsperelem, connectivity, points, centroid):
        self._lib.centroid(int(ndim), int(nelem), int(npointsperelem), self._p(connectivity, np.int64),
                           self._p(points, np.float64), self._p(centroid, np.float64))

    def triLinearInterpolator(self, nelem_to_search, npoints, nearest_element_indices, connectivity,
                              enclosing_elem_indices, nodes, weights, points):
        return int(self._lib.triLinearInterpolator(
            int(nelem_to_search), int(npoints), self._p(nearest_element_indices, np.int64),
            self._p(connectivity, np.int64), self._p(enclosing_elem_indices, np.int64),
            self._p(nodes, np.float64), self._p(weights, np.float64), self._p(points, np.float64)))

    def __getattr__(self, name):
        return getattr(self._lib, name)


def load_lib():
    if cache:
        return cache[0]
    lib = _NumpyLib(_lib.load_lib())
    cache.append(lib)
    return lib
