import csv, sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
seq=[(r[ki].split('(')[0].replace('void ','').replace('<unnamed>::',''), float(r[vi].replace(',',''))) for r in rows[1:]]
# last fused step = last occurrence of query_rank_kernel ... interp_coherent_kernel
idx=[i for i,(n,_) in enumerate(seq) if n.startswith('interp_coherent_kernel')]
end=idx[-1]; start=max(i for i,(n,_) in enumerate(seq[:end]) if n.startswith('query_rank_kernel') or n.startswith('histogram_kernel'))
tot=0
for n,v in seq[start:end+1]:
    print(f"{n[:64]:64s} {v/1e3:10.1f} us"); tot+=v
print(f"{'fused step total':64s} {tot/1e3:10.1f} us")
