#!/bin/bash
# im2col sequencing of the convs in the training step; side workloads with roofline / cpu_baseline
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_fusion.py -q -x -s 2>&1 | tail -n 8
timeout 900 python bench.py --workload train_b32 > gpurun_out/r2ae_train_n1.json 2> gpurun_out/r2ae_train_n1.err; echo "train rc=$?"; tail -n 3 gpurun_out/r2ae_train_n1.err
python tools/show_bench.py gpurun_out/r2ae_train_n1.json | head -16
for w in baseline_b16 latency_480; do timeout 600 python bench.py --workload $w > gpurun_out/r2ae_$w.json 2> gpurun_out/r2ae_$w.err; echo "$w rc=$?"; tail -n 3 gpurun_out/r2ae_$w.err; python - <<PY
import json
d=json.loads(open("gpurun_out/r2ae_$w.json").read().strip().splitlines()[-1])
print(d["value"], d["unit"], d["ms_per_step"], d["config"].get("eager_ms"), d["roofline"], d.get("cpu_baseline"))
PY
done
