#!/bin/bash
# quick GPU check: tensor-core unit tests first (bounded), then the whole GPU suite, then a short bench
set -u
TAG=${1:-q}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -x -q > gpurun_out/${TAG}_pytest_tc.log 2>&1; echo "tc exit $?" >> gpurun_out/${TAG}_pytest_tc.log
tail -15 gpurun_out/${TAG}_pytest_tc.log
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_tc.py > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "gpu exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -15 gpurun_out/${TAG}_pytest_gpu.log
for cfg in ${CFGS:-"0,-1"}; do
  bn=${cfg%,*}; pair=${cfg#*,}
  MAGE_TC_BN=$bn MAGE_TC_PAIR=$pair timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/${TAG}_bench_${bn}_${pair}.json 2> gpurun_out/${TAG}_bench_${bn}_${pair}.err
  echo "== bn=$bn pair=$pair"; python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_bench_${bn}_${pair}.json"))
    r = d["roofline"]
    print(d["value"], d["ms_per_step"], "gemm", r["achieved"], r["ms_in_kernel_per_step"], "conv", r["conv_implicit_gemm"], {k: v["ms_per_step"] for k, v in r["breakdown_ms_per_step"].items()})
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/${TAG}_bench_${bn}_${pair}.err").read()[-2000:])
PY
done
