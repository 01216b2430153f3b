#!/bin/bash
# the whole GPU suite + smoke() + a short default bench line, as the driver runs them
TAG=${1:-r02al}
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests/ -x -q -m gpu ) > gpurun_out/${TAG}_pytest_gpu_all.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu_all.log
tail -6 gpurun_out/${TAG}_pytest_gpu_all.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench.json')); print('bench', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'kps', d['kernels_per_step'], 'parity', d['parity']['token_mismatches'], 'eager', d['eager_gpu_baseline']['value'], 'cpu', d['cpu_baseline']['value'], 'frac', d['roofline']['frac'], d['clocks'])"
