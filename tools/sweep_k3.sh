for cfg in "4 2 8" "8 2 6" "8 2 4" "6 2 8" "8 2 8" "8 3 4"; do
set -- $cfg
MM_COH_SLOTS=$1 MM_COH_STAGES=$2 MM_COH_WARPS=$3 timeout -s KILL 100 python bench.py --workload S2 --steps 5 --warmup 3 --no-cpu 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('slots,stages,warps=$cfg', 'step ms', round(d['ms_per_step'],2), 'K3', d['kernels']['K3_interp']['ms'])
"
done
