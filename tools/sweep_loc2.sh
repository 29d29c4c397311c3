for v in 0 1 2 3; do for wl in S5small S2; do
MM_LOC_VARIANT=$v timeout -s KILL 100 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('variant $v $wl', 'step ms', round(d['ms_per_step'],3), 'K2', d['kernels']['K2_locate']['ms'])
"
done; done
