#!/bin/bash
# r02b: full GPU test suite on the refactored engine, then the small-batch schedule sweep (chunk streams x decode group) and the
# new default bench lines at B=64 / B=8.
set -u
TAG=${1:-r02b}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
grep "\[parity\]" gpurun_out/${TAG}_pytest_gpu.log | cut -c1-400 | head -60
tail -5 gpurun_out/${TAG}_pytest_gpu.log
run() {  # name, batch, env...
  local name=$1; local B=$2; shift; shift
  env "$@" timeout 600 python bench.py --batch $B --steps 10 --warmup 3 --no-cpu --no-parity --eager-gpu 0 > gpurun_out/${TAG}_${name}.json 2> gpurun_out/${TAG}_${name}.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_${name}.json"))
    print("${name}", d["value"], "f/s", d["ms_per_step"], "ms; e2e", d["e2e"]["value"], d["config"]["chunk_streams"], d["config"]["decode_group_frames"], d["kernels_per_step"])
except Exception as e:
    print("${name} failed", e); print(open("gpurun_out/${TAG}_${name}.err").read()[-1500:])
PY
}
run b8_s1_g4 8 MAGE_STREAMS=1 MAGE_DECODE_GROUP=4
run b8_s2_g8 8 MAGE_STREAMS=2 MAGE_DECODE_GROUP=8
run b8_s2_g16 8 MAGE_STREAMS=2 MAGE_DECODE_GROUP=16
run b8_s4_g8 8 MAGE_STREAMS=4 MAGE_DECODE_GROUP=8
run b8_s4_g16 8 MAGE_STREAMS=4 MAGE_DECODE_GROUP=16
run b8_s8_g16 8 MAGE_STREAMS=8 MAGE_DECODE_GROUP=16
run b8_s4_g16_pdl 8 MAGE_STREAMS=4 MAGE_DECODE_GROUP=16 MAGE_PDL=1
run b16_s1_g4 16 MAGE_STREAMS=1 MAGE_DECODE_GROUP=4
run b16_s2_g8 16 MAGE_STREAMS=2 MAGE_DECODE_GROUP=8
run b16_s4_g8 16 MAGE_STREAMS=4 MAGE_DECODE_GROUP=8
run b32_s1_g4 32 MAGE_STREAMS=1 MAGE_DECODE_GROUP=4
run b32_s2_g4 32 MAGE_STREAMS=2 MAGE_DECODE_GROUP=4
run b32_s4_g4 32 MAGE_STREAMS=4 MAGE_DECODE_GROUP=4
run b64_s1_g4 64 MAGE_STREAMS=1 MAGE_DECODE_GROUP=4
run b64_s2_g4 64 MAGE_STREAMS=2 MAGE_DECODE_GROUP=4
run b64_s4_g4 64 MAGE_STREAMS=4 MAGE_DECODE_GROUP=4
# the default line (with parity, eager GPU, cpu baseline) and the reference arm
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 2500 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_bench_reference.json | cut -c1-900
