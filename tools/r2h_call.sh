#!/bin/bash
# training slice: GPU tests of the training kernels + the trainer, then the train_b32 bench line at N = 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train.py tests/test_gpu_trainer.py -q -x -rxXs 2>&1 | tail -n 25
timeout 600 python bench.py --workload train_b32 > gpurun_out/r2h_train_n1.json 2> gpurun_out/r2h_train_n1.err; echo "bench rc=$?"
tail -n 5 gpurun_out/r2h_train_n1.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2h_train_n1.json").read())
print(round(d["ms_per_step"],3), "ms/step", round(d["value"]), "fps  e2e", round(d["e2e"]["value"]), "launches", d["gpu_launches"])
for k,v in list(d["kernels"].items())[:14]: print(f'{k:32s} n={v["launches_per_step"]:5.1f} {v["ms_per_step"]:.3f} ms  {v.get("bound")} {v.get("roofline_frac")}')
print(d["roofline"])
PY
