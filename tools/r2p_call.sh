#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -rxXs > gpurun_out/r2p_gpu_tests.log 2>&1; echo "suite rc=$?"; tail -n 3 gpurun_out/r2p_gpu_tests.log
timeout 300 python bench.py --workload tail_b16 > gpurun_out/r2p_tail.json 2> gpurun_out/r2p_tail.err; echo "tail rc=$?"; tail -n 3 gpurun_out/r2p_tail.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2p_tail.json").read())
print(round(d["ms_per_step"],3), "ms/step", round(d["value"]), "fps  e2e", round(d["e2e"]["value"]), "launches", d["gpu_launches"])
for k,v in list(d["kernels"].items())[:16]: print(f'{k:32s} n={v["launches_per_step"]:5.1f} {v["ms_per_step"]:.3f} ms  {v.get("bound")} {v.get("roofline_frac")}')
PY
