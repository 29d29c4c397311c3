#!/usr/bin/env python
"""Per-source-line instruction counts of one kernel from an .ncu-rep (read here, no GPU needed).

ncu's CSV source page lists SASS only; the line table comes from `nvdisasm --print-line-info` of the kernel's
cubin (extracted from the object file), matched to the SASS rows by instruction order.
usage: python tools/ncu_lines.py <rep> <object.o> <kernel-regex> <mangled-substring> [min-share]
"""
import csv
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict


def main():
    rep, obj, kregex, mangled = sys.argv[1:5]
    min_share = float(sys.argv[5]) if len(sys.argv) > 5 else 0.01
    sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", f"regex:{kregex}"],
                          capture_output=True, text=True).stdout
    rows = list(csv.reader(sass.splitlines()))
    h = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
    hdr = rows[h]
    ci, si, sm = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
    ti = hdr.index("Thread Instructions Executed")
    body = []
    for r in rows[h + 1:]:
        if len(r) > ci and r[ci].isdigit():
            body.append((r[si].strip(), int(r[ci]), int(r[sm]) if r[sm].isdigit() else 0, int(r[ti])))
        elif r and r[0] == "Kernel Name":
            break  # next kernel instance
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
    cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout
    lines = dis.splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith("\t.section\t.text.") and mangled in l)
    cur = None
    table = []  # per instruction: (file, line)
    for l in lines[start + 1:]:
        if l.startswith("\t.section"):
            break
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
            table.append(cur)
    n = min(len(table), len(body))
    if len(table) != len(body):
        print(f"# warning: {len(table)} instructions in the disassembly, {len(body)} in the report", file=sys.stderr)
    agg = defaultdict(lambda: [0, 0, 0])
    total = sum(b[1] for b in body)
    for i in range(n):
        a = agg[table[i]]
        a[0] += body[i][1]
        a[1] += body[i][2]
        a[2] += body[i][3]
    src_cache = {}
    print(f"total warp instructions {total}")
    for key, (cnt, samples, tcnt) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        if cnt < total * min_share:
            continue
        text = ""
        if key:
            path = os.path.join(os.path.dirname(os.path.abspath(obj)), "..", "..", "csrc", key[0])
            if key[0] not in src_cache and os.path.exists(path):
                src_cache[key[0]] = open(path).read().splitlines()
            if key[0] in src_cache and key[1] <= len(src_cache[key[0]]):
                text = src_cache[key[0]][key[1] - 1].strip()[:100]
        print(f"{100 * cnt / total:5.1f}%  lanes {tcnt / max(cnt, 1):4.1f}  samples {samples:6d}  {key}  {text}")


if __name__ == "__main__":
    main()
