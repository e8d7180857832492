#!/bin/bash
# last GSA epilogue writes NCHW; TMA conv default at C >= 64: parity, bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "parity or layers or depth or shard" > gpurun_out/r2y_tests.log 2>&1; echo "tests rc=$?"; tail -n 12 gpurun_out/r2y_tests.log
run() { tag=$1; shift; env "$@" timeout 240 python bench.py --no-cpu > gpurun_out/r2y_$tag.json 2> gpurun_out/r2y_$tag.err; python tools/show_bench.py gpurun_out/r2y_$tag.json 2>/dev/null | grep -v "^cpu\|^roofline" | head -24; tail -n 3 gpurun_out/r2y_$tag.err; }
run nchw CFP_GRAPH=1
