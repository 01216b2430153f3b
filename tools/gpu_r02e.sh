#!/bin/bash
# r02e: opaque-handle ABI, resident-weight halo convolutions: full GPU suite, conv microbench 3-pass / 1-pass, bench lines
set -u
TAG=${1:-r02e}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_ops.py -x -q > gpurun_out/${TAG}_pytest_tc.log 2>&1; echo "tc exit $?" >> gpurun_out/${TAG}_pytest_tc.log
tail -6 gpurun_out/${TAG}_pytest_tc.log
python tools/tc_microbench.py --only "dec " --passes 3 > gpurun_out/${TAG}_micro_p3.txt 2>&1; cat gpurun_out/${TAG}_micro_p3.txt
python tools/tc_microbench.py --only "pixel" --passes 3 >> gpurun_out/${TAG}_micro_p3.txt 2>&1; tail -1 gpurun_out/${TAG}_micro_p3.txt
python tools/tc_microbench.py --only "128x128" --passes 1 > gpurun_out/${TAG}_micro_p1.txt 2>&1; cat gpurun_out/${TAG}_micro_p1.txt
timeout 1500 python -m pytest tests -m gpu -x -q -s --deselect tests/test_gpu_tc.py --deselect tests/test_gpu_ops.py > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -6 gpurun_out/${TAG}_pytest_gpu.log | cut -c1-400
run() {
  local name=$1; shift
  env "$@" > gpurun_out/${TAG}_${name}.json 2> gpurun_out/${TAG}_${name}.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_${name}.json"))
    r = d["roofline"]
    print("${name}", d["value"], "f/s", d["ms_per_step"], "ms; e2e", d["e2e"]["value"], "gemm", r["achieved"], "conv", r["conv_implicit_gemm"]["achieved"], {k: v["ms_per_step"] for k, v in list(r["breakdown_ms_per_step"].items())[:6]})
except Exception as e:
    print("${name} failed", e); print(open("gpurun_out/${TAG}_${name}.err").read()[-1500:])
PY
}
Q="--no-cpu --no-parity --eager-gpu 0 --steps 10"
run b64 timeout 600 python bench.py $Q
run b64_fp32 MAGE_DECODER_PRECISION=fp32 timeout 600 python bench.py $Q
run b8 timeout 600 python bench.py --batch 8 $Q
run c3 timeout 600 python bench.py --workload c3 $Q
timeout 900 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_memcheck.log 2>&1; tail -4 gpurun_out/${TAG}_memcheck.log
