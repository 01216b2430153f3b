#!/bin/bash
# final validation as the driver runs it: the whole GPU suite in one go, smoke(), the default bench line, the reference arm
set -u
TAG=${1:-r02z}
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests/ -x -q -m gpu ) > gpurun_out/${TAG}_pytest_gpu_all.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu_all.log
tail -6 gpurun_out/${TAG}_pytest_gpu_all.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/${TAG}_bench_reference.json 2>/dev/null; cut -c1-260 gpurun_out/${TAG}_bench_reference.json
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench.json')); print('bench', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'parity', d['parity']['token_mismatches'], 'eager', d['eager_gpu_baseline']['value'], 'cpu', d['cpu_baseline']['value'], 'frac', d['roofline']['frac'], d['clocks'])"
Q="--no-cpu --no-parity --eager-gpu 0 --steps 10"
for i in 1 2 3; do
  for pdl in 0 1; do
    MAGE_PDL=$pdl python bench.py --batch 8 $Q 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('b8 pdl=$pdl', d['value'], d['ms_per_step'])"
  done
done
for pdl in 0 1; do
  MAGE_PDL=$pdl python bench.py --batch 16 $Q 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('b16 pdl=$pdl', d['value'], d['ms_per_step'])"
  MAGE_PDL=$pdl python bench.py $Q 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('b64 pdl=$pdl', d['value'], d['ms_per_step'])"
done
