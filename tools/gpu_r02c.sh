#!/bin/bash
# r02c: f4 VQ-VAE on tcgen05, decoder precision budget, small-problem tile model (BN=192 pair tiles)
set -u
TAG=${1:-r02c}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py -x -q > gpurun_out/${TAG}_pytest_tc.log 2>&1; echo "tc exit $?" >> gpurun_out/${TAG}_pytest_tc.log
tail -8 gpurun_out/${TAG}_pytest_tc.log
timeout 1500 python -m pytest tests -m gpu -x -q -s --deselect tests/test_gpu_tc.py > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
grep "\[parity\] decoder" gpurun_out/${TAG}_pytest_gpu.log | cut -c1-300
tail -8 gpurun_out/${TAG}_pytest_gpu.log | cut -c1-600
run() {  # name, args..., env via leading VAR=val handled by env
  local name=$1; shift
  env "$@" > gpurun_out/${TAG}_${name}.json 2> gpurun_out/${TAG}_${name}.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_${name}.json"))
    r = d["roofline"]
    print("${name}", d["value"], "f/s", d["ms_per_step"], "ms; e2e", d["e2e"]["value"], "gemm", r["achieved"], "conv", r["conv_implicit_gemm"]["achieved"], {k: v["ms_per_step"] for k, v in r["breakdown_ms_per_step"].items()})
except Exception as e:
    print("${name} failed", e); print(open("gpurun_out/${TAG}_${name}.err").read()[-1500:])
PY
}
Q="--no-cpu --no-parity --eager-gpu 0"
run b64_budget MAGE_DECODER_PRECISION=budget timeout 600 python bench.py --steps 10 $Q
run b64_fp32 MAGE_DECODER_PRECISION=fp32 timeout 600 python bench.py --steps 10 $Q
run b8_small1 MAGE_TC_SMALL=1 timeout 600 python bench.py --batch 8 --steps 10 $Q
run b8_small0 MAGE_TC_SMALL=0 timeout 600 python bench.py --batch 8 --steps 10 $Q
run b8_small1_pdl MAGE_TC_SMALL=1 MAGE_PDL=1 timeout 600 python bench.py --batch 8 --steps 10 $Q
run b16_small1 MAGE_TC_SMALL=1 timeout 600 python bench.py --batch 16 --steps 10 $Q
run b16_small0 MAGE_TC_SMALL=0 timeout 600 python bench.py --batch 16 --steps 10 $Q
run b32 timeout 600 python bench.py --batch 32 --steps 10 $Q
run c4 timeout 600 python bench.py --workload c4 --steps 10 $Q
run c3 timeout 600 python bench.py --workload c3 --steps 10 $Q
run c3_simt MAGE_BACKEND=simt timeout 600 python bench.py --workload c3 --steps 5 $Q
run c2 timeout 600 python bench.py --workload c2 --steps 10 $Q
