"""One fused step (query sort .. K3) out of an `ncu --metrics gpu__time_duration.sum` launch list of bench.py:
python tools/show_launches.py profiles/r2_S2_launches.csv [which interp_elem launch, default 2]"""
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
seq = [(r[ki].split('(')[0].replace('void ', '').replace('<unnamed>::', ''), float(r[vi].replace(',', '')))
       for r in rows[1:]]
ends = [i for i, (n, _) in enumerate(seq) if n.startswith('interp_elem_kernel')]
end = ends[int(sys.argv[2]) if len(sys.argv) > 2 else 2]
start = max(i for i, (n, _) in enumerate(seq[:end]) if n.startswith('query_rank_kernel'))
tot = 0
for n, v in seq[start:end + 1]:
    print(f"{n[:72]:72s} {v / 1e3:10.1f} us")
    tot += v
print(f"{'fused step total':72s} {tot / 1e3:10.1f} us")
