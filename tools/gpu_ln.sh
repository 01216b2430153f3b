#!/bin/bash
# fused LayerNorm (mage_gemm_tc_ln / mage_token_taps_ln_f32): bit-exactness first, then same-box A/B, then the whole GPU suite
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -q -x -k "fused_layernorm" 2>&1 | tail -5
timeout 300 python -m pytest tests/test_gpu_ops.py -q -x -k "token_taps" 2>&1 | tail -3
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "optional_schedules" 2>&1 | tail -5
Q="--no-cpu --no-parity --eager-gpu 0 --steps 10"
run() { local name=$1; local b=$2; shift; shift; env "$@" timeout 600 python bench.py --batch $b $Q 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$name', 'B=$b', d['value'], d['ms_per_step'], d['kernels_per_step'])"; }
{
for rep in 1 2 3; do
  run fused 64 MAGE_FUSED_LN=all
  run separate 64 MAGE_FUSED_LN=0
  run fused 8 MAGE_FUSED_LN=all
  run separate 8 MAGE_FUSED_LN=0
done
run fused 16 MAGE_FUSED_LN=all
run separate 16 MAGE_FUSED_LN=0
run fused 32 MAGE_FUSED_LN=all
run separate 32 MAGE_FUSED_LN=0
} 2>&1 | tee gpurun_out/ln_ab.txt
( time timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -15; echo "pytest exit ${PIPESTATUS[0]}" ) > gpurun_out/pytest_gpu_ln.log 2>&1
tail -8 gpurun_out/pytest_gpu_ln.log
