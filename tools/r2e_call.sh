#!/bin/bash
# programmatic dependent launch on every bf16-path kernel: parity suite, then A/B bench lines (PDL on / off, levels concurrent / sequential)
mkdir -p gpurun_out
timeout 800 python -m pytest tests -m gpu -q -x -rxXs > gpurun_out/r2e_gpu_tests.log 2>&1; echo "suite rc=$?"
tail -n 4 gpurun_out/r2e_gpu_tests.log
for pdl in on off; do for seq in 0 1; do
  tag="pdl_${pdl}_seq${seq}"
  env $( [ $pdl = off ] && echo CFP_NO_PDL=1 ) $( [ $seq = 1 ] && echo CFP_SEQUENTIAL_LEVELS=1 ) timeout 300 python bench.py --no-cpu > gpurun_out/r2e_bench_$tag.json 2> gpurun_out/r2e_bench_$tag.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2e_bench_$tag.json").read())
    print("$tag", round(d["ms_per_step"],3), "ms/step", round(d["value"]), "fps  e2e", round(d["e2e"]["value"]))
except Exception as e:
    print("$tag", "failed", e)
PY
done; done
