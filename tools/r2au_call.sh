#!/bin/bash
# N GPUs of one box, final build: headline line (device-resident + e2e with the host-link probe)
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus $N --no-cpu --steps 20 > gpurun_out/r2au_bench_n$N.json 2> gpurun_out/r2au_bench_n$N.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r2au_bench_n$N.json").read().strip().splitlines()[-1])
print("n$N", round(d["ms_per_step"],3), "ms/step", round(d["value"]), d["unit"], "e2e", round(d["e2e"]["value"]), "ceiling", round(d["e2e"]["host_link"]["ceiling_frames_per_s"]), d["e2e"]["host_link"]["GBps_each_way_per_gpu"])
PY
