#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/depth_attrib.py 2>&1 | tail -n 14 | tee gpurun_out/r2n_depth_attrib.txt
timeout 300 python bench.py --dtype f32 --no-cpu --steps 5 > gpurun_out/r2n_bench_f32.json 2> gpurun_out/r2n_bench_f32.err; echo rc=$?
python - <<PY
import json
d=json.loads(open("gpurun_out/r2n_bench_f32.json").read())
print("f32 engine", round(d["ms_per_step"],3), "ms/step", round(d["value"]), "fps  e2e", round(d["e2e"]["value"]))
PY
