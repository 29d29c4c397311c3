#!/usr/bin/env python
"""SASS evidence for profiles/: per kernel of the hot path, the static instruction mix that shows what it is built
from (UBLKCP = cp.async.bulk / TMA 1-D, SYNCS = mbarrier, DFMA/DADD/DMUL = fp64 pipe, FFMA + VIMNMX = the fp32
pre-filter of K1, MATCH/VOTE/SHFL = warp cooperation).   usage: python tools/sass_evidence.py > profiles/r2_sass_evidence.md
"""
import os
import re
import subprocess
from collections import Counter

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "multimesh_b200", "lib", "multi_mesh_b200.so")
COLS = ["UBLKCP", "SYNCS", "DFMA", "DADD", "DMUL", "FFMA", "FADD", "FMUL", "VIMNMX", "MATCH", "VOTE", "SHFL", "LDS", "STS",
        "LDG", "STG", "ATOMS", "ATOMG", "RED", "BAR"]
KEEP = re.compile(r"knn_tile_kernel|interp_elem|elem_rank|elem_place|key_rank|key_place|knn_kernel|knn_sites|locate_kernel|interp_tile|interp_coherent|interp_kernel|"
                  r"trilinear_kernel|trilinear_rec_kernel|gather_nodal|radix_scatter|scatter_back|query_place|element_geometry|fluid_fixup")
EXTRACT = {"interp_elem_kernel<2, 3>": r"UBLKCP|SYNCS|LDS.128", "interp_tile_kernel<2, 3>": r"UBLKCP|SYNCS", "locate_kernel<2, 3, 4, 8, 4, false, true, false>": r"UBLKCP|SYNCS|MATCH",
           "knn_tile_kernel<true, 4, 8>": r"VIMNMX|LDS.128|FFMA"}


def main():
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
    kernels, name = {}, None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            kernels[name] = []
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(.*?);", line)
        if m and name:
            kernels[name].append(m.group(1).strip())
    names = list(kernels)
    dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    pretty = {n: re.sub(r"\(anonymous namespace\)::", "", d).split("(")[0].replace("void ", "") for n, d in zip(names, dem)}
    print(f"# SASS evidence (`cuobjdump -sass multimesh_b200/lib/multi_mesh_b200.so`, sm_100a)\n")
    print("Static instruction counts per kernel. UBLKCP = `cp.async.bulk` (TMA 1-D bulk copy), SYNCS = mbarrier "
          "arrive / try_wait, DFMA/DADD/DMUL = binary64 pipe, FFMA/FADD/FMUL + VIMNMX = the fp32 pre-filter and key "
          "network of K1, MATCH/VOTE/SHFL = warp-level de-duplication and reductions.\n")
    print("| kernel | instr | " + " | ".join(COLS) + " |")
    print("|---|---|" + "---|" * len(COLS))
    for n in names:
        if not KEEP.search(pretty[n]):
            continue
        c = Counter()
        for ins in kernels[n]:
            op = re.sub(r"^@!?U?P\d+\s+", "", ins).split()[0].split(".")[0]
            c[op] += 1
        c["FADD"] += c.pop("FADD2", 0)
        print(f"| `{pretty[n]}` | {len(kernels[n])} | " + " | ".join(str(c.get(k, 0)) for k in COLS) + " |")
    for want, pat in EXTRACT.items():
        for n in names:
            if pretty[n] == want:
                print(f"\n### `{want}`: first matching instructions ({pat})\n\n```")
                hits = [i for i in kernels[n] if re.search(pat, i)]
                print("\n".join(hits[:16]))
                print("```")


if __name__ == "__main__":
    main()
