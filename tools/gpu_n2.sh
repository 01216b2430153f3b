#!/bin/bash
# forward half of the stage-2 objective + the other new tests of this stretch
mkdir -p gpurun_out
{
timeout 600 python -m pytest tests/test_gpu_ops.py -q -x -k "objective_kernels or conv3d_as_one" 2>&1 | tail -15
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -s -k "forward_loss" 2>&1 | grep -v "^$" | tail -60
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "operand_range" 2>&1 | tail -15
} 2>&1 | tee gpurun_out/n2_tests.log
