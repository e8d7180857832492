#!/bin/bash
# two GPUs: headline bench, training step (real 25 MB gradient bucket over NVLink), product sharding test
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $T bench.py --gpus 2 --no-cpu > gpurun_out/r2ac_bench_n2.json 2> gpurun_out/r2ac_bench_n2.err; echo "bench n2 rc=$?"
timeout 900 $T bench.py --gpus 2 --workload train_b32 > gpurun_out/r2ac_train_n2.json 2> gpurun_out/r2ac_train_n2.err; echo "train n2 rc=$?"; tail -n 3 gpurun_out/r2ac_train_n2.err
python - <<'PY'
import json
for f in ("gpurun_out/r2ac_bench_n2.json","gpurun_out/r2ac_train_n2.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d["value"],1), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), d.get("collective"))
    except Exception as e: print(f, "failed", e)
PY
timeout 600 python -m pytest tests/test_gpu_shard.py -q 2>&1 | tail -n 3
