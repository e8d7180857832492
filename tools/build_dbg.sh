#!/bin/bash
# debug variant of the library with the per-stage %globaltimer marks compiled in (k_chain_tc.cu, k_conv_tc.cu):
# cfpnet_b200/libcfp_dbg.so; use with CFP_LIB_PATH=cfpnet_b200/libcfp_dbg.so python tools/prof_layer.py ...
set -e
cd "$(dirname "$0")/.."
python -m cfpnet_b200.build
FL="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --expt-extended-lambda --expt-relaxed-constexpr -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -DCFP_DEBUG_TIMING"
for f in k_chain_tc k_conv_tc; do /usr/local/cuda/bin/nvcc $FL -c cfpnet_b200/csrc/$f.cu -o cfpnet_b200/build/${f}_dbg.o & done
wait
OBJS=$(ls cfpnet_b200/build/*.o | grep -v _dbg.o | grep -v -e k_chain_tc.o -e k_conv_tc.o)
/usr/local/cuda/bin/nvcc -shared -o cfpnet_b200/libcfp_dbg.so $OBJS cfpnet_b200/build/k_chain_tc_dbg.o cfpnet_b200/build/k_conv_tc_dbg.o -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -lcudart_static -lpthread -ldl -lrt
ls -la cfpnet_b200/libcfp_dbg.so
