#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "parity or layers or depth" 2>&1 | tail -n 5
timeout 300 python bench.py --no-cpu --steps 20 > gpurun_out/r2aj_bench.json 2> gpurun_out/r2aj_bench.err; python tools/show_bench.py gpurun_out/r2aj_bench.json 2>/dev/null | grep "value\|kv_state_tc<hist\|loftr_query_tc<\|attn_query_tc<dapm,128"
