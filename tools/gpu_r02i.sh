#!/bin/bash
# r02i: fused QKV + axial attention kernel: unit tests, full suite, bench A/B on one box, N=2 NCCL log check is separate
set -u
TAG=${1:-r02i}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py -x -q -k "fused_qkv" > gpurun_out/${TAG}_pytest_fused.log 2>&1; echo "fused exit $?" >> gpurun_out/${TAG}_pytest_fused.log
tail -25 gpurun_out/${TAG}_pytest_fused.log | cut -c1-250
timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_tc.py > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -8 gpurun_out/${TAG}_pytest_gpu.log | cut -c1-400
run() {
  local name=$1; shift
  env "$@" > gpurun_out/${TAG}_${name}.json 2> gpurun_out/${TAG}_${name}.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_${name}.json"))
    r = d["roofline"]
    print("${name}", d["value"], "f/s", d["ms_per_step"], "ms; e2e", d["e2e"]["value"], "k/step", d["kernels_per_step"], {k: v["ms_per_step"] for k, v in list(r["breakdown_ms_per_step"].items())[:6]})
except Exception as e:
    print("${name} failed", e); print(open("gpurun_out/${TAG}_${name}.err").read()[-1500:])
PY
}
Q="--no-cpu --no-parity --eager-gpu 0 --steps 10"
run b64_fused timeout 600 python bench.py $Q
run b64_unfused MAGE_FUSED_AXIAL=0 timeout 600 python bench.py $Q
run b64_fused2 timeout 600 python bench.py $Q
run b8_fused timeout 600 python bench.py --batch 8 $Q
run b8_unfused MAGE_FUSED_AXIAL=0 timeout 600 python bench.py --batch 8 $Q
