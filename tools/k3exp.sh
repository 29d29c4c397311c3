for cfg in "2 3 2" "1 3 4" "1 2 8" "5 3 2" "5 2 2" "5 2 3" "3 3 2" "2 2 4" "2 3 4"; do
set -- $cfg
MM_INTERP_FC=$1 MM_INTERP_STAGES=$2 MM_INTERP_WARPS=$3 python bench.py --workload medium --steps 3 --warmup 3 --no-cpu 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('FC,S,W=$cfg', 'K3 ms', d['kernels']['K3_interp']['ms'], 'frac', d['kernels']['K3_interp']['frac'], 'K2', d['kernels']['K2_locate']['ms'], 'K1', d['kernels']['K1_knn']['ms'])
"
done
