#!/bin/bash
# N = 8: headline line with the host-link probe (e2e ceiling per rank, all ranks at once), training step with the NCCL all-reduce
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2an_topo_n8.txt 2>&1
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus $N --no-cpu --steps 20 > gpurun_out/r2an_bench_n$N.json 2> gpurun_out/r2an_bench_n$N.err; echo "bench rc=$?"
timeout 600 $TR bench.py --gpus $N --workload train_b32 > gpurun_out/r2an_train_n$N.json 2> gpurun_out/r2an_train_n$N.err; echo "train rc=$?"
python - <<PY
import json
for f in ("gpurun_out/r2an_bench_n$N.json", "gpurun_out/r2an_train_n$N.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"],3), "ms/step", round(d["value"]), d["unit"], "e2e", d["e2e"], d.get("collective"))
    except Exception as e: print(f, "failed", e)
PY
tail -n 3 gpurun_out/r2an_bench_n$N.err gpurun_out/r2an_train_n$N.err
