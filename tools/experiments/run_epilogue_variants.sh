#!/bin/bash
# run on the GPU box: times the GEMM shapes of a decode step with each epilogue variant (see build_epilogue_variants.sh)
for v in BASE NO_RES NO_BIAS NO_STORE NO_FENCE BASE; do
  if [ $v = BASE ]; then unset MAGE_LIB; else export MAGE_LIB=$PWD/tools/experiments/libmage_exp_${v}.so; fi
  echo "== $v  (M=16384)"
  python tools/tc_microbench.py --iters 20 --only "x" 2>&1 | grep -v "^dec\|pixel\|conv3x3\|mainloop\|shape"
  echo "== $v  (M=2048)"
  python tools/tc_microbench.py --iters 20 --rows 2048 --only "x" 2>&1 | grep -v "^dec\|pixel\|conv3x3\|mainloop\|shape"
done
