#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "parity or layers or depth or lkpm" 2>&1 | tail -n 3
timeout 300 python bench.py --no-cpu --steps 20 > gpurun_out/r2al_bench.json 2> gpurun_out/r2al_bench.err
python tools/show_bench.py gpurun_out/r2al_bench.json 2>/dev/null | grep -v "^roofline\|^cpu" | cut -c1-110 | head -48
CFP_DW_FRAME=0 timeout 300 python bench.py --no-cpu --steps 20 > gpurun_out/r2al_bench_noframe.json 2>> gpurun_out/r2al_bench.err
python tools/show_bench.py gpurun_out/r2al_bench_noframe.json 2>/dev/null | grep "ms_per_step\|dwconv<7>" | cut -c1-100
export CFP_LIB_PATH=$PWD/cfpnet_b200/libcfp_dbg.so
for l in d2i dapm twins; do timeout 120 python tools/prof_layer.py $l 3 2>&1 | grep -v "^done" | tail -n 8; done > gpurun_out/r2al_timelines.txt 2>&1
sort -u gpurun_out/r2al_timelines.txt | cut -c1-330
