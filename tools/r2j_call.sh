#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 150 $TR bench.py --gpus 2 --workload train_b32 > gpurun_out/r2i_train_n2.json 2> gpurun_out/r2i_train_n2.err; echo "train rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r2i_train_n2.json").read().strip().splitlines()[-1])
print("train n2", round(d["ms_per_step"],3), "ms/step", round(d["value"]), "fps  e2e", round(d["e2e"]["value"]), d["collective"])
PY
tail -n 3 gpurun_out/r2i_train_n2.err
