#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "parity or layers or lkpm" 2>&1 | tail -n 3
timeout 300 python bench.py --no-cpu --steps 20 > gpurun_out/r2az_bench.json 2> gpurun_out/r2az_bench.err
python tools/show_bench.py gpurun_out/r2az_bench.json 2>/dev/null | grep "ms_per_step\|dwconv_tc" | cut -c1-150
