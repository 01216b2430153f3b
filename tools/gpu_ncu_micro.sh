#!/bin/bash
# ncu --set full on selected microbench shapes; exports raw/source CSV on the box.
#   tools/gpu_ncu_micro.sh TAG "shape substring" "cfgs" [env...]
set -u
TAG=$1; ONLY=$2; CFGS=${3:-"0,-1"}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tc_gemm_kernel|tc_conv_halo_kernel" -s 1 -c ${NCU_COUNT:-2} \
   -o gpurun_out/${TAG} python tools/tc_microbench.py --iters 1 --no-flush --only "$ONLY" --cfgs "$CFGS" > gpurun_out/${TAG}_log.txt 2>&1
ncu -i gpurun_out/${TAG}.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}.ncu-rep --page details --csv > gpurun_out/${TAG}_details.csv 2>/dev/null
ncu -i gpurun_out/${TAG}.ncu-rep --page source --csv --kernel-id :::1 > gpurun_out/${TAG}_source.csv 2>/dev/null
sz=$(stat -c %s gpurun_out/${TAG}.ncu-rep); if [ "$sz" -gt 20000000 ]; then rm -f gpurun_out/${TAG}.ncu-rep; fi
tail -3 gpurun_out/${TAG}_log.txt
