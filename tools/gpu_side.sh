#!/bin/bash
# small-batch schedule: VQ-VAE decoder beside the decode steps on a share of the SMs (MAGE_SIDE_SMS / _FRAMES / _GROUP)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "optional_schedules" 2>&1 | tail -5
Q="--no-cpu --no-parity --eager-gpu 0 --steps 10"
run() { local name=$1; local b=$2; shift; shift; env "$@" timeout 600 python bench.py --batch $b $Q 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$name', 'B=$b', d['value'], d['ms_per_step'], d['kernels_per_step'])"; }
{
run base 8 A=1
for D in 16 20 28 40; do for F in 8 12 16; do
  run "D=$D,F=$F,G=2" 8 MAGE_SIDE_SMS=$D MAGE_SIDE_FRAMES=$F MAGE_SIDE_GROUP=2
done; done
run base 8 A=1
run "D=20,F=12,G=4" 8 MAGE_SIDE_SMS=20 MAGE_SIDE_FRAMES=12 MAGE_SIDE_GROUP=4
run "D=20,F=12,G=1" 8 MAGE_SIDE_SMS=20 MAGE_SIDE_FRAMES=12 MAGE_SIDE_GROUP=1
run "D=28,F=20,G=4" 8 MAGE_SIDE_SMS=28 MAGE_SIDE_FRAMES=20 MAGE_SIDE_GROUP=4
run "D=40,F=24,G=4" 8 MAGE_SIDE_SMS=40 MAGE_SIDE_FRAMES=24 MAGE_SIDE_GROUP=4
run base 16 A=1
run "D=20,F=8,G=2" 16 MAGE_SIDE_SMS=20 MAGE_SIDE_FRAMES=8 MAGE_SIDE_GROUP=2
run "D=28,F=12,G=2" 16 MAGE_SIDE_SMS=28 MAGE_SIDE_FRAMES=12 MAGE_SIDE_GROUP=2
} 2>&1 | tee gpurun_out/side_ab.txt
