#!/bin/bash
mkdir -p gpurun_out
CFP_BENCH_DEBUG=1 timeout 300 python bench.py --no-cpu --steps 20 > gpurun_out/r2ai_bench.json 2> gpurun_out/r2ai_bench.err
grep "e2e debug" gpurun_out/r2ai_bench.err
python tools/show_bench.py gpurun_out/r2ai_bench.json 2>/dev/null | head -1
