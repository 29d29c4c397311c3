for v in 0 1 2 3; do
for wl in medium S5small; do
MM_LOC_VARIANT=$v timeout -s KILL 100 python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('variant $v $wl', 'step ms', round(d['ms_per_step'],3), 'K2 ms', d['kernels']['K2_locate']['ms'], 'K1', d['kernels']['K1_knn']['ms'], 'K3', d['kernels']['K3_interp']['ms'])
"
done; done
