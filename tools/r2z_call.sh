#!/bin/bash
# k = 7 depthwise conv (1/16 scale) on the Toeplitz tensor-core kernel vs the CUDA-core stencil
mkdir -p gpurun_out
CFP_DW_TC_MIN=7 timeout 600 python -m pytest tests -m gpu -q -x -k "parity or layers or depth" > gpurun_out/r2z_tests.log 2>&1; echo "tests rc=$?"; tail -n 12 gpurun_out/r2z_tests.log
run() { tag=$1; shift; env "$@" timeout 240 python bench.py --no-cpu > gpurun_out/r2z_$tag.json 2> gpurun_out/r2z_$tag.err; python tools/show_bench.py gpurun_out/r2z_$tag.json 2>/dev/null | grep "value\|dwconv\|dw_plane\|lkpm_mlp"; tail -n 3 gpurun_out/r2z_$tag.err; }
run dw7tc CFP_GRAPH=1 CFP_DW_TC_MIN=7
run dw7cc CFP_GRAPH=1
