#!/bin/bash
export CFP_LIB_PATH=$PWD/cfpnet_b200/libcfp_dbg.so
for lv in 1 2 3; do echo "== level $lv"; timeout 120 python tools/prof_layer.py dapm $lv 2>&1 | grep "conv3x3" | sort | uniq -c | sort -rn | head -6; done
