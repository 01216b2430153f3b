#!/bin/bash
# r02d: MAGE+ branch tests, TC unit tests, single-pass conv microbench + ncu, chunk streams with the new tile model
set -u
TAG=${1:-r02d}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py -x -q > gpurun_out/${TAG}_pytest_tc.log 2>&1; echo "tc exit $?" >> gpurun_out/${TAG}_pytest_tc.log
tail -4 gpurun_out/${TAG}_pytest_tc.log
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -s -k "mage_plus or precision" > gpurun_out/${TAG}_pytest_plus.log 2>&1; echo "plus exit $?" >> gpurun_out/${TAG}_pytest_plus.log
grep "\[parity\]" gpurun_out/${TAG}_pytest_plus.log | cut -c1-300; tail -15 gpurun_out/${TAG}_pytest_plus.log | cut -c1-300
python tools/tc_microbench.py --only "128x128" --passes 3 > gpurun_out/${TAG}_micro_p3.txt 2>&1; cat gpurun_out/${TAG}_micro_p3.txt
python tools/tc_microbench.py --only "128x128" --passes 1 > gpurun_out/${TAG}_micro_p1.txt 2>&1; cat gpurun_out/${TAG}_micro_p1.txt
python tools/tc_microbench.py --rows 2048 --only "x" > gpurun_out/${TAG}_micro_m2048.txt 2>&1; grep -v "^dec\|pixel\|conv3x3\|mainloop" gpurun_out/${TAG}_micro_m2048.txt
MAGE_TC_SMALL=0 python tools/tc_microbench.py --rows 2048 --only "x" > gpurun_out/${TAG}_micro_m2048_small0.txt 2>&1; grep -v "^dec\|pixel\|conv3x3\|mainloop" gpurun_out/${TAG}_micro_m2048_small0.txt
# ncu --set full: pixel head and dec 128x128 64->64 with 1 pass
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tc_conv_halo_kernel" -s 2 -c 2 \
   -o gpurun_out/${TAG}_halo_p1 python tools/tc_microbench.py --iters 1 --no-flush --only "128x128" --passes 1 > gpurun_out/${TAG}_ncu_p1.log 2>&1
ncu -i gpurun_out/${TAG}_halo_p1.ncu-rep --page raw --csv > gpurun_out/${TAG}_halo_p1_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_halo_p1.ncu-rep --page details --csv > gpurun_out/${TAG}_halo_p1_details.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_halo_p1.ncu-rep --page source --csv --kernel-id :::1 > gpurun_out/${TAG}_halo_p1_source.csv 2>/dev/null
python tools/ncu_summary.py raw gpurun_out/${TAG}_halo_p1_raw.csv gpurun_out/${TAG}_halo_p1_summary.csv
run() {
  local name=$1; shift
  env "$@" > gpurun_out/${TAG}_${name}.json 2> gpurun_out/${TAG}_${name}.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_${name}.json"))
    print("${name}", d["value"], "f/s", d["ms_per_step"], "ms; e2e", d["e2e"]["value"], d["config"]["chunk_streams"], d["config"]["decode_group_frames"])
except Exception as e:
    print("${name} failed", e); print(open("gpurun_out/${TAG}_${name}.err").read()[-1500:])
PY
}
Q="--no-cpu --no-parity --eager-gpu 0 --steps 10"
run b8_s1 MAGE_STREAMS=1 timeout 600 python bench.py --batch 8 $Q
run b8_s2 MAGE_STREAMS=2 timeout 600 python bench.py --batch 8 $Q
run b8_s2_g8 MAGE_STREAMS=2 MAGE_DECODE_GROUP=8 timeout 600 python bench.py --batch 8 $Q
run b8_s1_ov MAGE_STREAMS=1 MAGE_OVERLAP_DECODE=1 MAGE_DECODE_GROUP=8 timeout 600 python bench.py --batch 8 $Q
run b16_s2 MAGE_STREAMS=2 timeout 600 python bench.py --batch 16 $Q
