"""
Build recipes for the CPU oracle (TEST INFRASTRUCTURE ONLY -- see oracle/mm_oracle.c header).

  build_oracle()  gcc-compiles oracle/mm_oracle.c  -> oracle/_build/libmm_oracle.so
  build_ref()     gcc-compiles the reference's own C sources, where they lie under
                  /root/reference/multi_mesh/src (centroid.c, trilinearinterpolator.c), with the
                  flags the reference intended (setup.py:14-15: -O3 -fopenmp)
                  -> oracle/_ref/multi_mesh_ref.so.  Sources are never copied into this repo.
                  oracle/_ref/ is git-ignored but travels to the GPU box with the snapshot.

Run as a script:  python oracle/build.py [--ref]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD_DIR = os.path.join(HERE, "_build")
REF_DIR = os.path.join(HERE, "_ref")
ORACLE_SO = os.path.join(BUILD_DIR, "libmm_oracle.so")
REF_SO = os.path.join(REF_DIR, "multi_mesh_ref.so")
REF_SRC = "/root/reference/multi_mesh/src"


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources if os.path.exists(s))


def build_oracle(force=False):
    src = os.path.join(HERE, "mm_oracle.c")
    if not force and not _stale(ORACLE_SO, [src, __file__]):
        return ORACLE_SO
    os.makedirs(BUILD_DIR, exist_ok=True)
    cmd = [
        "gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-fopenmp",
        # canonical arithmetic: one rounding per operation, no contraction, no fast-math
        "-ffp-contract=off", "-fno-fast-math", "-fexcess-precision=standard",
        "-Wall", "-Wextra", "-o", ORACLE_SO, src, "-lm",
    ]
    subprocess.check_call(cmd)
    return ORACLE_SO


def ref_available():
    return os.path.exists(os.path.join(REF_SRC, "trilinearinterpolator.c"))


def build_ref(force=False):
    """Compile the reference's C path from its own sources (only possible where
    /root/reference exists, i.e. in the build container; the GPU box uses the prebuilt .so)."""
    if not ref_available():
        return REF_SO if os.path.exists(REF_SO) else None
    srcs = [os.path.join(REF_SRC, "centroid.c"), os.path.join(REF_SRC, "trilinearinterpolator.c")]
    if not force and not _stale(REF_SO, srcs + [__file__]):
        return REF_SO
    os.makedirs(REF_DIR, exist_ok=True)
    cmd = ["gcc", "-O3", "-fopenmp", "-fPIC", "-shared", "-o", REF_SO] + srcs + ["-lm", "-lgomp"]
    subprocess.check_call(cmd)
    return REF_SO


if __name__ == "__main__":
    print(build_oracle(force=True))
    if "--ref" in sys.argv or ref_available():
        print(build_ref(force=True))
