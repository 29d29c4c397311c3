for cfg in "5 2 3 8" "1 8 3 8" "1 4 3 8" "2 4 3 8" "1 8 2 8" "5 4 2 8" "1 16 2 8"; do
set -- $cfg
MM_COH_FC=$1 MM_COH_SLOTS=$2 MM_COH_STAGES=$3 MM_COH_WARPS=$4 timeout -s KILL 100 python bench.py --workload S2o4 --steps 2 --warmup 3 --no-cpu 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('FC,slots,stages,warps=$cfg', 'step ms', round(d['ms_per_step'],2), 'Mpts/s', round(d['value']/1e6,1))
"
done
