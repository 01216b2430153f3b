#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -s -k "forward_loss_mage_plus or (forward_loss_vs and forward_L4_b2) or mage_plus_branch_vs" 2>&1 | grep "parity\] forward\|passed\|failed\|Error\|error\|assert" | tail -30
timeout 600 python -m pytest tests/test_gpu_tc.py -q -x -k "gemm_tc_shapes and auto" 2>&1 | tail -2
