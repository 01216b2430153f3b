#!/bin/bash
# TIMING EXPERIMENT ONLY: builds variants of libmage_sm100.so with one piece of the tensor-core epilogue removed each (results are
# WRONG by construction), to measure what each piece costs (tools/experiments/run_epilogue_variants.sh, MAGE_LIB selects the library).
set -e
cd "$(dirname "$0")/../../mage_b200/csrc"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC --expt-relaxed-constexpr"
for v in NO_RES NO_BIAS NO_STORE NO_FENCE; do
  out=../../tools/experiments/libmage_exp_${v}.so
  $NVCC $FLAGS -DMAGE_EXP_${v} -shared -o $out misc.cu gemm_simt.cu gemm_tc.cu attention.cu vq.cu -lcuda
  echo built $out
done
