#!/bin/bash
mkdir -p gpurun_out
for v in "CFP_CONV_TCOLS=256" "CFP_CONV_TMA=1" "CFP_CONV_TMA=0" "CFP_DW_TC_MIN=7"; do
  env $v timeout 300 python bench.py --no-cpu --steps 20 > gpurun_out/r2ba_bench.json 2> gpurun_out/r2ba_bench.err
  echo "== $v"; python tools/show_bench.py gpurun_out/r2ba_bench.json 2>/dev/null | grep "ms_per_step" | cut -c1-80
done
