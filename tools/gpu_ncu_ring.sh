#!/bin/bash
# ncu --set full of the c_proj GEMM at 8 prompts (M = 2048: ring-bound main loop) and of c_fc at 64 prompts (M = 16384: MMA-bound)
set -u
mkdir -p gpurun_out
for spec in "r02aq_ncu_proj_m2048 2048 proj" "r02aq_ncu_fc_m16384 16384 fc"; do
  set -- $spec; TAG=$1; ROWS=$2; ONLY=$3
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tc_gemm_kernel" -s 1 -c 1 \
     -o gpurun_out/${TAG} python tools/tc_microbench.py --rows $ROWS --iters 2 --no-flush --only "$ONLY $ROWS" --cfgs "0,-1" > gpurun_out/${TAG}_log.txt 2>&1
  ncu -i gpurun_out/${TAG}.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>/dev/null
  ncu -i gpurun_out/${TAG}.ncu-rep --page source --csv > gpurun_out/${TAG}_source.csv 2>/dev/null
  python tools/ncu_summary.py raw gpurun_out/${TAG}_raw.csv gpurun_out/${TAG}_summary.csv | head -5
  python tools/ncu_hot_sass.py gpurun_out/${TAG}_source.csv 30 > gpurun_out/${TAG}_hot_sass.txt 2>&1; head -3 gpurun_out/${TAG}_hot_sass.txt
  rm -f gpurun_out/${TAG}.ncu-rep gpurun_out/${TAG}_source.csv
  tail -2 gpurun_out/${TAG}_log.txt
done
