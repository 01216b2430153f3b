#!/bin/bash
# strong-scaling lines at N = 8 and N = 4 on one 8-GPU box (torchrun, the driver's launch line)
mkdir -p gpurun_out
for N in ${NS:-8 4}; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29540 + N)) bench.py --gpus $N --steps 5 --warmup 3 \
      > gpurun_out/${TAG:-r02n}_bench_n$N.json 2> gpurun_out/${TAG:-r02n}_bench_n$N.err
  python - <<PY
import json
try:
    txt = open("gpurun_out/${TAG:-r02n}_bench_n$N.json").read().strip().splitlines()
    print("stdout lines:", len(txt))
    d = json.loads(txt[-1])
    print("N=$N", d["value"], "f/s", d["ms_per_step"], "ms", d["scaling"], "e2e", d["e2e"]["value"], "weak", d.get("weak_scaling"), "parity", d["parity"]["token_mismatches"], d["parity"]["oracle_seconds"], d["config"]["batch_per_gpu"])
except Exception as e:
    print("N=$N failed", e); print(open("gpurun_out/${TAG:-r02n}_bench_n$N.err").read()[-2000:])
PY
  grep -c "NCCL INFO" gpurun_out/${TAG:-r02n}_bench_n$N.err; grep "NCCL INFO" gpurun_out/${TAG:-r02n}_bench_n$N.err | grep -i "nranks" | head -2 | cut -c1-200
done
