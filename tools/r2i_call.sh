#!/bin/bash
# N = 2: trainer / shard GPU tests over NCCL, train_b32 line, headline line with the host-link probe
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2i_topo.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_trainer.py tests/test_gpu_shard.py -q -x -rxXs 2>&1 | tail -n 8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 2 --workload train_b32 > gpurun_out/r2i_train_n2.json 2> gpurun_out/r2i_train_n2.err; echo "train rc=$?"
timeout 600 $TR bench.py --gpus 2 --no-cpu > gpurun_out/r2i_bench_n2.json 2> gpurun_out/r2i_bench_n2.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r2i_train_n2.json").read().strip().splitlines()[-1])
print("train n2", round(d["ms_per_step"],3), "ms/step", round(d["value"]), "fps  e2e", round(d["e2e"]["value"]), d["collective"])
d=json.loads(open("gpurun_out/r2i_bench_n2.json").read().strip().splitlines()[-1])
print("bench n2", round(d["ms_per_step"],3), "ms/step", round(d["value"]), "fps  e2e", d["e2e"])
PY
tail -n 3 gpurun_out/r2i_train_n2.err gpurun_out/r2i_bench_n2.err
