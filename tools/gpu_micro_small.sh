#!/bin/bash
# per-shape tile sweep of the decode-step GEMMs at 8 and 16 prompts (M = 2048 / 4096 rows), warm L2 (operands of a step are L2-resident).
# The first column is a clock warm-up (its numbers are discarded); the automatic choice (bn0,pair-1) is measured twice, 2nd and last column.
mkdir -p gpurun_out
{
for R in 2048 4096; do
echo "## rows $R (warm cache)"
python tools/tc_microbench.py --rows $R --no-flush --iters 20 --only "x" --cfgs "128,1;0,-1;256,1;192,1;128,1;64,1;256,0;128,0;64,0;0,-1" 2>&1 | grep -v "mainloop\|16x16\|conv3x3"
done
} | tee gpurun_out/r02ak_tc_microbench_small_m.txt
