#!/bin/bash
# decode-group sweep at the headline batch (frames decoded per VQ-VAE decoder pass), one box
mkdir -p gpurun_out
Q="--no-cpu --no-parity --eager-gpu 0 --steps 8"
run() { local name=$1; local b=$2; shift; shift; env "$@" timeout 600 python bench.py --batch $b $Q 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$name', 'B=$b', d['value'], d['ms_per_step'], d['kernels_per_step'])"; }
{
for rep in 1 2; do
for G in 4 2 8 16 31; do run "G=$G" 64 MAGE_DECODE_GROUP=$G; done
done
run "G=4,overlap" 64 MAGE_DECODE_GROUP=4 MAGE_OVERLAP_DECODE=1
run "G=8,overlap" 64 MAGE_DECODE_GROUP=8 MAGE_OVERLAP_DECODE=1
} 2>&1 | tee gpurun_out/group_sweep_b64.txt
