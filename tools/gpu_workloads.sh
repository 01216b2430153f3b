#!/bin/bash
# the other BASELINE configs + the MAGE+ branch + the validation workload on the final state (short lines, no CPU / eager legs)
mkdir -p gpurun_out
Q="--no-cpu --eager-gpu 0 --steps 5"
for W in c2 c3 c4 c5plus c4val; do
  timeout 900 python bench.py --workload $W $Q > gpurun_out/r02ar_bench_$W.json 2> gpurun_out/r02ar_bench_$W.err
  python -c "
import json; d=json.loads(open('gpurun_out/r02ar_bench_$W.json').read().strip().splitlines()[-1]); print('$W', d['value'], d['unit'], d['ms_per_step'], 'ms e2e', d['e2e']['value'], 'kps', d['kernels_per_step'], 'parity', {k: v for k, v in (d.get('parity') or {}).items() if k in ('token_mismatches', 'latent_max_abs_err', 'rel_err')})"
done
