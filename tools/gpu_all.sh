#!/bin/bash
# whole GPU suite + default bench lines, logs into gpurun_out/
mkdir -p gpurun_out
TAG=${1:-r02ac}
( time timeout 2400 python -m pytest tests/ -x -q -m gpu ) > gpurun_out/${TAG}_pytest_gpu_all.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu_all.log
tail -6 gpurun_out/${TAG}_pytest_gpu_all.log
grep -h "\[parity\] --split val\|\[parity\] forward" gpurun_out/${TAG}_pytest_gpu_all.log | head
