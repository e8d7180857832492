#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --workload train_b32 > gpurun_out/r2ab_train_n1.json 2> gpurun_out/r2ab_train_n1.err; echo "rc=$?"; tail -n 5 gpurun_out/r2ab_train_n1.err
python tools/show_bench.py gpurun_out/r2ab_train_n1.json | head -30
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2ab_train_n1.json").read().strip().splitlines()[-1]); print(d["collective"]); print(d["config"])
PY
nvidia-smi --query-gpu=memory.used --format=csv
