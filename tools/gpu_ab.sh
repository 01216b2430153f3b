#!/bin/bash
# A/B of the halo-convolution weight handling on ONE box: resident slots vs ring, 3-pass and single-pass
mkdir -p gpurun_out
for r in 1 0 1 0; do
  for p in 3 1; do
    echo "== resident=$r passes=$p"
    MAGE_TC_RESIDENT=$r python tools/tc_microbench.py --only "dec " --passes $p --iters 20 | tail -7
    MAGE_TC_RESIDENT=$r python tools/tc_microbench.py --only "pixel" --passes $p --iters 20 | tail -1
  done
done
