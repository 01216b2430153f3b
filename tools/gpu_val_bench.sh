#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_ops.py -q -x -s -k "forward_loss or split_val or temporal_attn_seq" 2>&1 | grep "parity\]\|passed\|failed\|Error\|error" | tail -30
timeout 900 python bench.py --workload c4val > gpurun_out/r02ae_bench_c4val.json 2> gpurun_out/r02ae_bench_c4val.err; tail -c 3000 gpurun_out/r02ae_bench_c4val.json; tail -5 gpurun_out/r02ae_bench_c4val.err
timeout 600 python bench.py --workload c4val --batch 64 --no-cpu --no-parity --eager-gpu 0 > gpurun_out/r02ae_bench_c4val_b64.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r02ae_bench_c4val_b64.json').read()); print('b64', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['breakdown_ms_per_step'])"
