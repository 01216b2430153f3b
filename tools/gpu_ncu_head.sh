#!/bin/bash
# ncu --set full + source view of the fused pixel-head convolution in single-pass and 3-pass mode
set -u
mkdir -p gpurun_out
for P in 1 3; do
  TAG=r02j_head_p${P}
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_conv_halo_kernel -s 1 -c 1 \
     -o gpurun_out/${TAG} python tools/tc_microbench.py --iters 1 --no-flush --only "pixel" --passes $P > gpurun_out/${TAG}_log.txt 2>&1
  ncu -i gpurun_out/${TAG}.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>/dev/null
  ncu -i gpurun_out/${TAG}.ncu-rep --page details --csv > gpurun_out/${TAG}_details.csv 2>/dev/null
  ncu -i gpurun_out/${TAG}.ncu-rep --page source --csv > gpurun_out/${TAG}_source.csv 2>/dev/null
  rm -f gpurun_out/${TAG}.ncu-rep
  python tools/ncu_summary.py raw gpurun_out/${TAG}_raw.csv gpurun_out/${TAG}_summary.csv | tail -2
  python tools/ncu_hot_sass.py gpurun_out/${TAG}_source.csv 45
done
ncu --query-metrics 2>/dev/null | grep -i "tensor\|tmem\|pipe_uniform\|utc" | head -40 > gpurun_out/r02j_tensor_metrics.txt; cat gpurun_out/r02j_tensor_metrics.txt
