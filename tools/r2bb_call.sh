#!/bin/bash
mkdir -p gpurun_out
for v in "CFP_X=0" "CFP_CONV_TMA=0" "CFP_CONV_TMA=0 CFP_DW_TC_MIN=7"; do
  env $v timeout 300 python bench.py --no-cpu --steps 20 > gpurun_out/r2bb_bench.json 2> gpurun_out/r2bb_bench.err
  echo "== $v"; python tools/show_bench.py gpurun_out/r2bb_bench.json 2>/dev/null | grep "ms_per_step\|conv3x3\|dwconv\|dw_plane\|lkpm_mlp_tc<128" | cut -c1-80
done
