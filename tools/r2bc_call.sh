#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "parity or layers or depth" 2>&1 | tail -n 3
timeout 300 python bench.py --no-cpu --steps 20 > gpurun_out/r2bc_bench.json 2> gpurun_out/r2bc_bench.err
python tools/show_bench.py gpurun_out/r2bc_bench.json 2>/dev/null | grep "ms_per_step" | cut -c1-100
CFP_SEQUENTIAL_LEVELS=1 timeout 300 python bench.py --no-cpu --steps 20 > gpurun_out/r2bc_bench_seq.json 2> gpurun_out/r2bc_bench.err
python tools/show_bench.py gpurun_out/r2bc_bench_seq.json 2>/dev/null | grep "ms_per_step" | cut -c1-100
