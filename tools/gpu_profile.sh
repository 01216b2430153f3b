#!/bin/bash
# One gpurun call: GPU tests, the default bench line, the ncu launch list of the bench command and
# full ncu captures of the dominant kernels.  Everything lands in gpurun_out/ (scratch); summaries
# are copied into profiles/ by tools/ncu_summary.py afterwards.
set -u
TAG=${1:-r01}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -3 gpurun_out/${TAG}_pytest_gpu.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -c 3000 gpurun_out/${TAG}_bench.json
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err
# launch list of the same bench command: one whole CUDA-graph replay of generate (skip the eager warm-up + first replay)
KPS=${KPS:-2494}
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s $((KPS * 2)) -c $KPS --csv \
    --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/${TAG}_ncu_bench.log 2>&1
# full capture: one decode step worth of tensor-core GEMM / implicit-GEMM launches, and two temporal-attention launches
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_kernel -s ${FULL_SKIP:-852} -c ${FULL_COUNT:-54} \
    -o gpurun_out/${TAG}_tc python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/${TAG}_ncu_tc.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:temporal_attn -s 40 -c 2 \
    -o gpurun_out/${TAG}_tattn python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/${TAG}_ncu_tattn.log 2>&1
ls -la gpurun_out
