#!/bin/bash
# Builds tools/experiments/libmage_oneacc.so: the product sources with -DMAGE_EXPERIMENT_ONEACC (see gemm_tc.cu: timing only,
# numerically wrong with the shipped split format).  Use:  MAGE_LIB=tools/experiments/libmage_oneacc.so python tools/tc_microbench.py ...
set -e
cd "$(dirname "$0")/../../mage_b200/csrc"
OUT=../../tools/experiments
NV="/usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC --expt-relaxed-constexpr"
mkdir -p /tmp/oneacc
for f in misc gemm_simt attention vq; do $NV -c $f.cu -o /tmp/oneacc/$f.o; done
$NV -DMAGE_EXPERIMENT_ONEACC -c gemm_tc.cu -o /tmp/oneacc/gemm_tc.o
$NV -shared -o $OUT/libmage_oneacc.so /tmp/oneacc/*.o -lcuda
ls -la $OUT/libmage_oneacc.so
