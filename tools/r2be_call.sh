#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do for v in "CFP_X=0" "CFP_DW_TC_MIN=7"; do
  env $v timeout 300 python bench.py --no-cpu --steps 20 > gpurun_out/r2be_bench.json 2> gpurun_out/r2be_bench.err
  echo "== $v: $(python tools/show_bench.py gpurun_out/r2be_bench.json 2>/dev/null | grep ms_per_step | cut -c1-70)"
  env $v CFP_SEQUENTIAL_LEVELS=1 timeout 300 python bench.py --no-cpu --steps 20 > gpurun_out/r2be_bench.json 2> gpurun_out/r2be_bench.err
  echo "   seq $v: $(python tools/show_bench.py gpurun_out/r2be_bench.json 2>/dev/null | grep ms_per_step | cut -c1-70)"
done; done
