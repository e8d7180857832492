#!/bin/bash
# micro-batch streams inside TransformerFusion.forward: parity test, then a sweep (levels sequential / concurrent)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_layers.py -q -x -k micro 2>&1 | tail -n 3
for seq in 1 0; do for micro in "128:1,64:1,32:1" "128:2,64:1,32:1" "128:2,64:2,32:1" "128:4,64:2,32:1" "128:4,64:2,32:2" "128:4,64:4,32:2" "128:8,64:4,32:2"; do
  tag="seq${seq}_$(echo $micro | tr ':,' '__')"
  env CFP_MICRO=$micro $( [ $seq = 1 ] && echo CFP_SEQUENTIAL_LEVELS=1 ) timeout 200 python bench.py --no-cpu > gpurun_out/r2k_$tag.json 2> gpurun_out/r2k_$tag.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2k_$tag.json").read())
    print("$tag", round(d["ms_per_step"],3), "ms/step", round(d["value"]), "fps  e2e", round(d["e2e"]["value"]))
except Exception as e:
    print("$tag", "failed", e)
PY
done; done
