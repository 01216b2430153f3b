#!/bin/bash
# One gpurun call: [tests] [bench] [launch list] [full ncu captures].  Everything lands in gpurun_out/ (scratch, <= 64 MiB:
# the .ncu-rep files are exported to CSV on the box and dropped when large); tools/ncu_summary.py condenses them into profiles/.
#   usage: tools/gpu_profile.sh TAG "phases"     phases in {test bench ref list full tattn}
set -u
TAG=${1:-r01}
PH=${2:-"test bench ref list full tattn"}
mkdir -p gpurun_out
has() { [[ " $PH " == *" $1 "* ]]; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
if has test; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1
  echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
  tail -5 gpurun_out/${TAG}_pytest_gpu.log
fi
if has bench; then
  timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
  tail -c 3500 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
fi
if has ref; then
  timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err
fi
KPS=${KPS:-2494}
if has list; then
  # launch list of the bench command: one whole CUDA-graph replay of generate (skips the eager warm-up + first replay)
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s $((KPS * 2)) -c $KPS --csv \
      --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/${TAG}_ncu_bench.log 2>&1
fi
export_rep() {  # $1 = report stem
  if [ -f gpurun_out/$1.ncu-rep ]; then
    ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null
    ncu -i gpurun_out/$1.ncu-rep --page details --csv > gpurun_out/$1_details.csv 2>/dev/null
    ncu -i gpurun_out/$1.ncu-rep --page source --csv --kernel-id :::1 > gpurun_out/$1_source_first.csv 2>/dev/null
    sz=$(stat -c %s gpurun_out/$1.ncu-rep)
    if [ "$sz" -gt 30000000 ]; then rm -f gpurun_out/$1.ncu-rep; fi
  fi
}
if has full; then
  # one decode step worth of tensor-core launches is 54; capture every ${FULL_EVERY}th of them around step 16
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s ${FULL_SKIP:-852} -c ${FULL_COUNT:-54} \
      -o gpurun_out/${TAG}_tc python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/${TAG}_ncu_tc.log 2>&1
  export_rep ${TAG}_tc
fi
if has tattn; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:temporal_attn -s 40 -c 2 \
      -o gpurun_out/${TAG}_tattn python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/${TAG}_ncu_tattn.log 2>&1
  export_rep ${TAG}_tattn
fi
du -sh gpurun_out; ls -la gpurun_out
