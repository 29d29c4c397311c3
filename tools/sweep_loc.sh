#!/bin/bash
# K2 launch-configuration variants (MM_LOC_VARIANT, profiling only) on the bench workload
WL=${WL:-S2}
for v in ${VARIANTS:-0 4 5 6 7}; do
  MM_LOC_VARIANT=$v timeout -s KILL 300 python bench.py --workload $WL --steps 5 --warmup 3 --no-cpu > gpurun_out/loc_$v.json 2> gpurun_out/loc_$v.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/loc_$v.json").read().strip().splitlines()[-1])
print("$WL variant $v", d["ms_per_step"], {k:v["ms"] for k,v in d["kernels"].items()})
PY
done
