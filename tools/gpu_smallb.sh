#!/bin/bash
# Small-batch (strong-scaling) regime: bench lines at B prompts per GPU under different schedules + an ncu launch list.
#   usage: tools/gpu_smallb.sh TAG [B]
set -u
TAG=${1:-r02a}
B=${2:-8}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
run() {  # name, env...
  local name=$1; shift
  env "$@" timeout 600 python bench.py --batch $B --steps 10 --warmup 3 --no-cpu > gpurun_out/${TAG}_b${B}_${name}.json 2> gpurun_out/${TAG}_b${B}_${name}.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_b${B}_${name}.json"))
    r = d["roofline"]
    print("${name}", d["value"], "f/s", d["ms_per_step"], "ms; e2e", d["e2e"]["value"], {k: (v["ms_per_step"], v["launches"]) for k, v in r["breakdown_ms_per_step"].items()})
except Exception as e:
    print("${name} failed", e); print(open("gpurun_out/${TAG}_b${B}_${name}.err").read()[-1500:])
PY
}
run base MAGE_PDL=0
run pdl MAGE_PDL=1
run group32 MAGE_PDL=0 MAGE_DECODE_GROUP=32
run group32_pdl MAGE_PDL=1 MAGE_DECODE_GROUP=32
run group32_overlap MAGE_PDL=0 MAGE_DECODE_GROUP=32 MAGE_OVERLAP_DECODE=1
KPS=${KPS:-1720}
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s $((KPS * 2)) -c $KPS --csv \
    --log-file gpurun_out/${TAG}_b${B}_launches.csv python bench.py --batch $B --steps 1 --warmup 3 --no-cpu > gpurun_out/${TAG}_b${B}_ncu.log 2>&1
python tools/ncu_summary.py launches gpurun_out/${TAG}_b${B}_launches.csv gpurun_out/${TAG}_b${B}_launches_summary.csv | head -40
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/${TAG}_b64.json 2> gpurun_out/${TAG}_b64.err
python -c "
import json; d=json.load(open('gpurun_out/${TAG}_b64.json')); print('b64', d['value'], d['ms_per_step'], d['e2e']['value'])"
