#!/bin/bash
# small-batch A/B on one box: single-CTA tiles only (no clusters), store wait at kernel end, PDL
Q="--no-cpu --no-parity --eager-gpu 0 --steps 10"
run() { local name=$1; shift; env "$@" python bench.py --batch 8 $Q 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$name', d['value'], d['ms_per_step'], d['kernels_per_step'])"; }
run base A=1
run nopair MAGE_TC_PAIR=0
run wait_read MAGE_LIB=$PWD/tools/experiments/libmage_exp_WAIT_READ.so
run base2 A=1
run small0 MAGE_TC_SMALL=0
run wait_read_pdl MAGE_LIB=$PWD/tools/experiments/libmage_exp_WAIT_READ.so MAGE_PDL=1
run g31 MAGE_DECODE_GROUP=31
run g8_overlap MAGE_DECODE_GROUP=8 MAGE_OVERLAP_DECODE=1
