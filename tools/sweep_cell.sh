#!/bin/bash
# K1 sensitivity to the index cell size (MM_INDEX_CELL_SCALE multiplies the auto-tuned cell)
WL=${WL:-S2}
for sc in ${SCALES:-1.0 1.2 1.5 2.0 0.8}; do
  MM_INDEX_CELL_SCALE=$sc timeout -s KILL 300 python bench.py --workload $WL --steps 5 --warmup 3 --no-cpu > gpurun_out/cell_$sc.json 2> gpurun_out/cell_$sc.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/cell_$sc.json").read().strip().splitlines()[-1])
print("$WL $sc", d["ms_per_step"], {k:v["ms"] for k,v in d["kernels"].items()}, d["other_stages"], d["index_build_ms"])
PY
done
