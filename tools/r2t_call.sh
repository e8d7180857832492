#!/bin/bash
# two threads per token row in the C >= 64 query chains (and kv_state at C = 64): parity, then A/B bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "parity or layers or depth or shard" > gpurun_out/r2t_tests.log 2>&1; echo "tests rc=$?"; tail -n 4 gpurun_out/r2t_tests.log
run() { tag=$1; shift; env "$@" timeout 240 python bench.py --no-cpu > gpurun_out/r2t_$tag.json 2> gpurun_out/r2t_$tag.err; python tools/show_bench.py gpurun_out/r2t_$tag.json 2>/dev/null | head -60; }
run nt2 CFP_GRAPH=1
run nt1 CFP_GRAPH=1 CFP_CHAIN_NT=1
