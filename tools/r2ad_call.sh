#!/bin/bash
# which round-2 change costs the small-batch workloads (baseline_b16 1.50 -> 1.65 ms, latency_480 0.81 -> 0.86 ms)?
mkdir -p gpurun_out
run() { tag=$1; w=$2; shift 2; env "$@" timeout 300 python bench.py --workload $w > gpurun_out/r2ad_${w}_$tag.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r2ad_${w}_$tag.json').read().strip().splitlines()[-1]); print('$w $tag', round(d['ms_per_step'],4))"; }
for w in baseline_b16 latency_480; do
run default $w A=1
run nt1 $w CFP_CHAIN_NT=1
run notma $w CFP_CONV_TMA=0
run nt1_notma $w CFP_CHAIN_NT=1 CFP_CONV_TMA=0
done
