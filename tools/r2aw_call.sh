#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "stream_host or graphed" 2>&1 | tail -n 5
for k in 10 20; do
timeout 300 python bench.py --no-cpu --steps $k > gpurun_out/r2aw_bench_k$k.json 2> gpurun_out/r2aw_bench.err
CFP_COARSE_EVENTS=1 timeout 300 python bench.py --no-cpu --steps $k > gpurun_out/r2aw_bench_coarse_k$k.json 2>> gpurun_out/r2aw_bench.err
python - <<PY
import json
for f in ("gpurun_out/r2aw_bench_k$k.json","gpurun_out/r2aw_bench_coarse_k$k.json"):
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d["ms_per_step"],3), round(d["value"]), "e2e", round(d["e2e"]["value"]), d.get("e2e_attempts_ms"))
PY
done
tail -3 gpurun_out/r2aw_bench.err
