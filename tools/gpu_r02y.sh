#!/bin/bash
set -u
TAG=r02y
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests/ -x -q -m gpu ) > gpurun_out/${TAG}_pytest_gpu_all.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu_all.log
tail -5 gpurun_out/${TAG}_pytest_gpu_all.log
Q="--no-cpu --no-parity --eager-gpu 0 --steps 10"
for i in 1 2; do
  for pdl in 0 1; do
    for B in 8 16 64; do
      MAGE_PDL=$pdl python bench.py --batch $B $Q 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('b$B pdl=$pdl', d['value'], d['ms_per_step'])"
    done
  done
done
