#!/bin/bash
# per-stage %globaltimer timelines (debug library) of the L3 / L2 chains, two threads per row vs one
export CFP_LIB_PATH=$PWD/cfpnet_b200/libcfp_dbg.so
for lv in 3 2; do
for nt in 0 1; do
  echo "=== level $lv CFP_CHAIN_NT=$nt"
  for l in d2i dapm twins; do CFP_CHAIN_NT=$nt timeout 120 python tools/prof_layer.py $l $lv 2>&1 | grep -v "^done" | tail -n 12; done
done
done
