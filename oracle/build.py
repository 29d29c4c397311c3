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


def _cpu_has_fma():
    try:
        with open("/proc/cpuinfo") as fh:
            for line in fh:
                if line.startswith("flags"):
                    return " fma " in (line + " ")
    except OSError:
        pass
    return False


def build_oracle(force=False):
    src = os.path.join(HERE, "mm_oracle.c")
    # the explicit fma() calls of the canonical arithmetic inline to one instruction with -mfma; without it
    # they go through libm (same results, slower).  The library built in the build container travels to the
    # GPU box, so the flag set is recorded and the library is rebuilt where the CPU does not match it.
    arch = ["-mfma"] if _cpu_has_fma() else []
    stamp = os.path.join(BUILD_DIR, "flags.txt")
    want = " ".join(arch)
    have = open(stamp).read() if os.path.exists(stamp) else None
    if not force and have == want and not _stale(ORACLE_SO, [src, __file__]):
        return ORACLE_SO
    os.makedirs(BUILD_DIR, exist_ok=True)
    cmd = [
        "gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-fopenmp",
        # canonical arithmetic: nothing is contracted or re-associated by the compiler; the only fused
        # operations are the explicit fma() calls in the source
        "-ffp-contract=off", "-fno-fast-math", "-fexcess-precision=standard",
    ] + arch + ["-Wall", "-Wextra", "-o", ORACLE_SO, src, "-lm"]
    subprocess.check_call(cmd)
    with open(stamp, "w") as fh:
        fh.write(want)
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
