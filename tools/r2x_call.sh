#!/bin/bash
# conv3x3 raster staging through the tensor-map TMA engine: parity, then A/B bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "parity or layers or depth" > gpurun_out/r2x_tests.log 2>&1; echo "tests rc=$?"; tail -n 12 gpurun_out/r2x_tests.log
run() { tag=$1; shift; env "$@" timeout 240 python bench.py --no-cpu > gpurun_out/r2x_$tag.json 2> gpurun_out/r2x_$tag.err; python tools/show_bench.py gpurun_out/r2x_$tag.json 2>/dev/null | grep -v "^cpu\|^roofline" | grep "value\|conv3x3"; tail -n 3 gpurun_out/r2x_$tag.err; }
run tma CFP_GRAPH=1
run cpasync CFP_GRAPH=1 CFP_CONV_TMA=0
