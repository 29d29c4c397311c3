"""
The C-ABI boundary without a GPU: the library builds for sm_100a, loads, and exports every
function include/multimesh_b200.h declares; calls that need a device fail LOUDLY (error code +
message), never silently and never through a CPU fallback.
"""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "multimesh_b200.h")


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"^\s*#.*$", "", src, flags=re.M)  # preprocessor lines
    names = re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{}]*\)\s*;", src)
    return sorted(set(n for n in names if n not in ("defined",)))


@pytest.fixture(scope="module")
def lib():
    from multimesh_b200 import build, _lib

    build.build()
    return _lib.load_lib()


def test_header_declares_expected_entry_points():
    names = _declared_functions()
    for must in ("mm_knn", "mm_locate", "mm_interp", "mm_index_create", "mm_element_geometry", "mm_trilinear",
                 "centroid", "triLinearInterpolator", "mm_interpolate_host"):
        assert must in names


def test_library_exports_every_declared_symbol(lib):
    from multimesh_b200 import _lib

    path = _lib.library_path()
    out = subprocess.run(["nm", "-D", "--defined-only", path], capture_output=True, text=True, check=True).stdout
    exported = set(line.split()[-1] for line in out.splitlines() if line.strip())
    declared = _declared_functions()
    missing = [n for n in declared if n not in exported]
    assert not missing, f"declared in the header but not exported: {missing}"
    # and the Python binding table covers the header exactly
    assert sorted(_lib.EXPORTED_SYMBOLS) == declared


def test_library_is_built_for_sm_100a(lib):
    from multimesh_b200 import _lib

    out = subprocess.run(["cuobjdump", "-lelf", _lib.library_path()], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert lib.mm_version() == 100


def test_argument_validation_without_device(lib):
    from multimesh_b200 import _lib

    prm = _lib.LocateParams()
    rc = lib.mm_interp(3, 3, 1, 1, None, 1, None, None, None, None)  # order 3 unsupported
    assert rc == -1 and b"order" in lib.mm_last_error()
    rc = lib.mm_locate(2, 5, 1, None, None, None, None, 1, None, 20, None, C.byref(prm), None, None, None, None, None)
    assert rc == -1 and b"dim" in lib.mm_last_error()
    h = C.c_void_p()
    rc = lib.mm_index_create(C.byref(h), 7, 10, None, None)
    assert rc == -1
    rc = lib.mm_knn(None, 1, None, 20, 1, None, None, None)
    assert rc == -1 and b"null index" in lib.mm_last_error()


def test_no_cpu_fallback(lib):
    """Without a CUDA device the compute entry points return MM_ERR_CUDA; the Python ops refuse
    CPU tensors."""
    import torch
    from multimesh_b200 import ops

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    x = np.zeros((4, 8, 3))
    out = np.zeros((4, 3))
    rc = lib.mm_element_geometry(1, 3, 4, x.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), None, None)
    assert rc == -2 and b"CUDA" in lib.mm_last_error()
    with pytest.raises(Exception):
        ops.element_geometry(torch.zeros((4, 8, 3), dtype=torch.float64))
    with pytest.raises(Exception):
        ops.interp(torch.zeros((1, 1, 8), dtype=torch.float64), torch.zeros(1, dtype=torch.int32),
                   torch.zeros((1, 3), dtype=torch.float64))


def test_product_does_not_import_the_oracle():
    """The oracle is test infrastructure: nothing under multimesh_b200/ or multi_mesh/ may import,
    link or execute it."""
    bad = []
    for pkg in ("multimesh_b200", "multi_mesh"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, pkg)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h")):
                    txt = open(os.path.join(dirpath, f), errors="ignore").read()
                    if re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M) or "mm_oracle" in txt and f.endswith(".py"):
                        bad.append(os.path.join(dirpath, f))
                    if re.search(r'#include\s+".*oracle', txt):
                        bad.append(os.path.join(dirpath, f))
    assert not bad, bad
