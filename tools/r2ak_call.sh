#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "parity or layers" 2>&1 | tail -n 3
for sp in 1 0; do for seq in 0 1; do
  if [ $seq = 1 ]; then export CFP_SEQUENTIAL_LEVELS=1; else unset CFP_SEQUENTIAL_LEVELS; fi
  CFP_CHAIN_SPREAD=$sp timeout 300 python bench.py --no-cpu --steps 20 > gpurun_out/r2ak_bench_sp${sp}_seq${seq}.json 2> gpurun_out/r2ak.err
  echo "spread=$sp seq=$seq"; python tools/show_bench.py gpurun_out/r2ak_bench_sp${sp}_seq${seq}.json 2>/dev/null | grep "ms_per_step\|loftr_query_tc<.*128\|attn_query_tc<dapm,128" | cut -c1-100
done; done
