#!/bin/bash
# persistent ring form of the temporal attention step: bit-exactness, then same-box A/B
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -q -x -k "temporal_attn" 2>&1 | tail -5
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "optional_schedules or (generate_vs_reference_golden and tc)" 2>&1 | tail -4
Q="--no-cpu --no-parity --eager-gpu 0 --steps 10"
run() { local name=$1; local b=$2; shift; shift; env "$@" timeout 600 python bench.py --batch $b $Q 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$name', 'B=$b', d['value'], d['ms_per_step'], d['kernels_per_step'], 'tattn', r.get('temporal_attention'))"; }
{
for rep in 1 2 3; do
  run ring 64 MAGE_TATTN_RING=1
  run oneshot 64 MAGE_TATTN_RING=0
  run ring 8 MAGE_TATTN_RING=1
  run oneshot 8 MAGE_TATTN_RING=0
done
run ring 16 MAGE_TATTN_RING=1
run oneshot 16 MAGE_TATTN_RING=0
} 2>&1 | tee gpurun_out/tattn_ring_ab.txt
