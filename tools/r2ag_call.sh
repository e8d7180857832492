#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "parity or layers or depth" 2>&1 | tail -n 3
for w in latency_480 baseline_b16; do timeout 600 python bench.py --workload $w --no-cpu > gpurun_out/r2ag_$w.json 2> gpurun_out/r2ag_$w.err; python - <<PY
import json
d=json.loads(open("gpurun_out/r2ag_$w.json").read().strip().splitlines()[-1])
print("$w", round(d["value"],3), d["unit"], round(d["ms_per_step"],4), d["config"].get("eager_ms"))
for k,v in list(d.get("kernels",{}).items())[:4]: print("   ", k, v.get("launches_per_step"), round(v["ms_per_step"],4))
PY
done
