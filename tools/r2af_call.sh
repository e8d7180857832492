#!/bin/bash
# refresh the side / training / tail lines with the gap-free profiler
mkdir -p gpurun_out
for w in baseline_b16 latency_480 train_b32 tail_b16; do timeout 900 python bench.py --workload $w > gpurun_out/r2af_$w.json 2> gpurun_out/r2af_$w.err; echo "$w rc=$?"; tail -n 2 gpurun_out/r2af_$w.err; python - <<PY
import json
d=json.loads(open("gpurun_out/r2af_$w.json").read().strip().splitlines()[-1])
print(round(d["value"],2), d["unit"], round(d["ms_per_step"],4), d["config"].get("eager_ms") if isinstance(d["config"],dict) else None, (d.get("roofline") or {}).get("kernel"), (d.get("roofline") or {}).get("frac"), d.get("cpu_baseline"))
for k,v in list(d.get("kernels",{}).items())[:6]: print("   ", k, v.get("launches_per_step"), round(v["ms_per_step"],4), v.get("roofline_frac"))
PY
done
timeout 300 python bench.py --no-cpu > gpurun_out/r2af_bench.json 2> gpurun_out/r2af_bench.err; python tools/show_bench.py gpurun_out/r2af_bench.json 2>/dev/null | head -8
