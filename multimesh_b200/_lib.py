"""
ctypes binding of the C-ABI library (include/multimesh_b200.h).

Mirrors the reference's loader (multi_mesh/helpers.py:29-84): the shared object is looked up as
`lib/multi_mesh*.so` next to this package, cached, and a ValueError is raised when it is missing.
There is no fallback of any kind: if the CUDA library cannot be loaded the product path fails.
"""
import ctypes as C
import glob
import os

LIB_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib")

MM_OK = 0
FB_FAIL, FB_MAGIC, FB_SNAP, FB_MINL1 = 0, 1, 2, 3
(ST_ACCEPTED, ST_FB_INSIDE_MAGIC, ST_FB_NEAR_OK, ST_FB_NEAR_MAGIC, ST_FB_NAN_MAGIC, ST_SNAPPED,
 ST_FAILED, ST_MINL1, ST_SNAP_NONE) = range(9)


class LocateParams(C.Structure):
    """mm_locate_params (include/multimesh_b200.h)."""
    _fields_ = [
        ("aabb_prefilter", C.c_int32),
        ("strict", C.c_int32),
        ("fallback", C.c_int32),
        ("reserved", C.c_int32),
        ("tol", C.c_double),
        ("snap_clip", C.c_double),
        ("magic_xi", C.c_double * 3),
    ]


class MultiMeshError(RuntimeError):
    pass


_cache = []

_vp = C.c_void_p
_i64 = C.c_int64
_int = C.c_int

_SIGNATURES = {
    # name: (restype, argtypes)
    "mm_version": (_int, []),
    "mm_last_error": (C.c_char_p, []),
    "mm_element_geometry": (_int, [_int, _int, _i64, _vp, _vp, _vp, _vp]),
    "mm_map_to_sphere": (_int, [_i64, _vp, _vp, C.c_double, _vp]),
    "mm_index_create": (_int, [C.POINTER(_vp), _int, _i64, _vp, _vp]),
    "mm_index_destroy": (_int, [_vp]),
    "mm_index_info": (_int, [_vp, C.POINTER(_i64 * 8), C.POINTER(C.c_double)]),
    "mm_knn": (_int, [_vp, _i64, _vp, _int, C.c_int32, _vp, _vp, _vp]),
    "mm_element_presolve": (_int, [_int, _int, _i64, _vp, _vp, _vp]),
    "mm_locate": (_int, [_int, _int, _i64, _vp, _vp, _vp, _vp, _i64, _vp, _int, _vp,
                         C.POINTER(LocateParams), _vp, _vp, _vp, _vp, _vp]),
    "mm_interp": (_int, [_int, _int, _i64, _int, _vp, _i64, _vp, _vp, _vp, _vp]),
    "mm_interp_perm": (_int, [_int, _int, _i64, _int, _vp, _i64, _vp, _vp, _vp, _vp, _vp]),
    "mm_interpolate_workspace_bytes": (C.c_size_t, [_vp, _int, _i64, _int]),
    "mm_interpolate": (_int, [_vp, C.c_int32, _int, _int, _i64, _vp, _vp, _vp, _vp, _int, _vp, _i64, _vp, _int,
                              C.POINTER(LocateParams), _vp, _vp, _vp, _vp, _vp, _vp, C.c_size_t, _vp]),
    "mm_coeffs": (_int, [_int, _int, _i64, _vp, _vp, _vp, _vp]),
    "mm_gather_coeffs": (_int, [_int, _i64, _int, _vp, _i64, _vp, _vp, _vp, _vp]),
    "mm_trilinear": (_int, [_i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "mm_centroid_conn": (_int, [_i64, _i64, _i64, _vp, _vp, _vp, _vp]),
    "mm_trilinear_indexed_workspace_bytes": (C.c_size_t, [_vp, _i64, _int]),
    "mm_trilinear_indexed": (_int, [_vp, _i64, _vp, _vp, _i64, _vp, _int, _vp, _vp, _vp, _vp, C.c_size_t, _vp]),
    "mm_gather_nodal": (_int, [_int, _i64, _vp, _i64, _vp, _vp, _vp, _vp]),
    "mm_interpolate_host": (_int, [_int, _int, _i64, _vp, _int, _vp, _i64, _vp, _int, _int,
                                   C.POINTER(LocateParams), _vp, _vp, _vp, C.POINTER(_i64)]),
    "mm_locate_set_stats": (_int, [_vp]),
    "mm_unique_points": (_int, [_int, _i64, _vp, C.POINTER(_i64), _vp, _vp, _vp]),
    "mm_scatter_back": (_int, [_i64, _int, _int, _vp, _vp, _vp, _vp]),
    "mm_fluid_fixup": (_int, [_i64, _int, _int, _vp, _vp, _vp, _int, _vp]),
    "mm_host_release": (_int, []),
    "mm_pool_trim": (_int, []),
    "mm_index_prepare_sites": (_int, [_vp, _vp]),
    "mm_source_create_host": (_int, [C.POINTER(_vp), _int, _int, _i64, _vp, _int, _vp, _int]),
    "mm_source_create_device": (_int, [C.POINTER(_vp), _int, _int, _i64, _vp, _int, _vp, _int, _vp]),
    "mm_source_set_fields_host": (_int, [_vp, _int, _vp]),
    "mm_source_destroy": (_int, [_vp]),
    "mm_source_info": (_int, [_vp, C.POINTER(_i64 * 8)]),
    "mm_source_index": (_vp, [_vp]),
    "mm_source_interpolate": (_int, [_vp, _i64, _vp, _int, C.POINTER(LocateParams), _vp, _vp, _vp, _vp, _vp, _vp]),
    "mm_source_interpolate_host": (_int, [_vp, _i64, _vp, _int, C.POINTER(LocateParams), _vp, _vp, _vp,
                                          C.POINTER(_i64)]),
    "mm_profile_create": (_int, [C.POINTER(_vp), _int]),
    "mm_profile_destroy": (_int, [_vp]),
    "mm_profile_begin": (_int, [_vp]),
    "mm_profile_end": (_int, []),
    "mm_profile_read": (_int, [_vp, C.POINTER(_int), _vp]),
    # legacy symbols with the reference's signatures (host pointers)
    "centroid": (None, [C.c_longlong, C.c_longlong, C.c_longlong, _vp, _vp, _vp]),
    "triLinearInterpolator": (C.c_longlong, [C.c_longlong, C.c_longlong, _vp, _vp, _vp, _vp, _vp, _vp]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)


def library_path():
    files = sorted(glob.glob(os.path.join(LIB_DIR, "multi_mesh*.so")))
    if not files:
        raise ValueError(
            "Could not find suitable MultiMesh shared library (expected "
            f"{LIB_DIR}/multi_mesh*.so). Build it with `python -m multimesh_b200.build`; "
            "there is no CPU fallback."
        )
    return files[0]


def load_lib():
    if _cache:
        return _cache[0]
    lib = C.CDLL(library_path())
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export the symbol
        fn.restype = res
        fn.argtypes = args
    _cache.append(lib)
    return lib


def check(rc, what=""):
    if rc != MM_OK:
        msg = load_lib().mm_last_error()
        raise MultiMeshError(f"{what} failed with code {rc}: {msg.decode() if msg else ''}")
