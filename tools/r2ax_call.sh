#!/bin/bash
# refresh of the headline lines and the GPU suite after the last host-side change (per-level events in stream_host)
mkdir -p gpurun_out
P=gpurun_out/r2g
timeout 900 python -m pytest tests -m gpu -q -rxXs > ${P}_gpu_tests.log 2>&1; echo "suite rc=$?"; tail -n 3 ${P}_gpu_tests.log
python __graft_entry__.py --smoke > ${P}_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 ${P}_smoke.log
timeout 600 python bench.py > ${P}_bench.json 2> ${P}_bench.err; echo "bench rc=$?"
CFP_SEQUENTIAL_LEVELS=1 timeout 300 python bench.py --no-cpu > ${P}_bench_seq.json 2> ${P}_bench_seq.err; echo "bench seq rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > ${P}_bench_reference.json 2> ${P}_bench_reference.err; echo "reference rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/r2g_bench.json","gpurun_out/r2g_bench_seq.json","gpurun_out/r2g_bench_reference.json"):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f.split("/")[-1], round(d.get("value",0),1), d.get("unit"), "ms/step", round(d.get("ms_per_step",0),3), "e2e", round((d.get("e2e") or {}).get("value",0),1), (d.get("roofline") or {}).get("kernel"), round((d.get("roofline") or {}).get("frac") or 0,3), d.get("cpu_baseline"))
PY
