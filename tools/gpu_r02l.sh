#!/bin/bash
set -u
TAG=r02l
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py -x -q 2>&1 | tail -3
for P in 3 1; do python tools/tc_microbench.py --iters 20 --only "pixel" --passes $P 2>&1 | tail -1; done
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "golden or precision or vqvae" 2>&1 | tail -3
python bench.py --no-cpu --no-parity --eager-gpu 0 --steps 10 > gpurun_out/${TAG}_b64.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/${TAG}_b64.json')); r=d['roofline']; print('b64', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], {k: v['ms_per_step'] for k, v in list(r['breakdown_ms_per_step'].items())[:6]})"
python bench.py --batch 8 --no-cpu --no-parity --eager-gpu 0 --steps 10 > gpurun_out/${TAG}_b8.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/${TAG}_b8.json')); print('b8', d['value'], d['ms_per_step'], d['kernels_per_step'])"
