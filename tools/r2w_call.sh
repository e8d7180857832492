#!/bin/bash
# mbarrier try_wait suspend hint + hist2image cold paths out of line: parity, bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "parity or layers or depth or umma or decoder" > gpurun_out/r2w_tests.log 2>&1; echo "tests rc=$?"; tail -n 4 gpurun_out/r2w_tests.log
run() { tag=$1; shift; env "$@" timeout 240 python bench.py --no-cpu > gpurun_out/r2w_$tag.json 2> gpurun_out/r2w_$tag.err; python tools/show_bench.py gpurun_out/r2w_$tag.json 2>/dev/null | head -50; }
run fast CFP_GRAPH=1
