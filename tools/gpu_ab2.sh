#!/bin/bash
# A/B on one box: residual prefetch one chunk ahead (default build) vs issued after the chunk (experiment build)
for v in BASE NO_RES_PREFETCH BASE NO_RES_PREFETCH; do
  if [ $v = BASE ]; then unset MAGE_LIB; else export MAGE_LIB=$PWD/tools/experiments/libmage_exp_${v}.so; fi
  echo "== $v"
  python tools/tc_microbench.py --iters 20 --only "x" 2>&1 | grep -v "shape\|mainloop\|16x16\|32x32"
  python tools/tc_microbench.py --iters 20 --only "pixel" --passes 1 2>&1 | tail -1
done
unset MAGE_LIB
python bench.py --no-cpu --no-parity --eager-gpu 0 --steps 10 > gpurun_out/r02k_b64.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r02k_b64.json')); print('b64 default', d['value'], d['ms_per_step'])"
MAGE_LIB=$PWD/tools/experiments/libmage_exp_NO_RES_PREFETCH.so python bench.py --no-cpu --no-parity --eager-gpu 0 --steps 10 > gpurun_out/r02k_b64_nopf.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r02k_b64_nopf.json')); print('b64 no-prefetch', d['value'], d['ms_per_step'])"
python bench.py --no-cpu --no-parity --eager-gpu 0 --steps 10 > gpurun_out/r02k_b64b.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r02k_b64b.json')); print('b64 default', d['value'], d['ms_per_step'])"
timeout 600 python -m pytest tests/test_gpu_tc.py -x -q 2>&1 | tail -2
