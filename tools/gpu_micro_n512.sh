#!/bin/bash
# the N = 512 GEMMs of a decode step at 64 prompts (M = 16384): every tile shape, cold (L2 flushed) and warm
mkdir -p gpurun_out
{
echo "## rows 16384, L2 flushed between iterations (first column = clock warm-up, automatic choice measured 2nd and last)"
python tools/tc_microbench.py --rows 16384 --iters 20 --only "x512x" --cfgs "128,1;0,-1;256,1;128,1;64,1;256,0;128,0;64,0;0,-1" 2>&1 | grep -v "mainloop"
echo "## rows 16384, warm"
python tools/tc_microbench.py --rows 16384 --iters 20 --no-flush --only "x512x" --cfgs "128,1;0,-1;256,1;128,1;64,1;256,0;128,0;64,0;0,-1" 2>&1 | grep -v "mainloop"
echo "## all step shapes, automatic choice, flushed"
python tools/tc_microbench.py --rows 16384 --iters 20 --cfgs "0,-1;0,-1" 2>&1
} | tee gpurun_out/r02an_tc_microbench_n512.txt
