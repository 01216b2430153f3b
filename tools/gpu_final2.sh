#!/bin/bash
# final validation as the driver runs it (whole GPU suite, smoke(), reference arm, default bench line) + the ncu launch list of the
# default bench command + compute-sanitizer memcheck over the kernels added since the last sanitizer run
set -u
TAG=${1:-r02af}
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests/ -x -q -m gpu ) > gpurun_out/${TAG}_pytest_gpu_all.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu_all.log
tail -6 gpurun_out/${TAG}_pytest_gpu_all.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/${TAG}_bench_reference.json 2>/dev/null; cut -c1-200 gpurun_out/${TAG}_bench_reference.json
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench.json')); print('bench', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'kps', d['kernels_per_step'], 'parity', d['parity']['token_mismatches'], 'eager', d['eager_gpu_baseline']['value'], 'cpu', d['cpu_baseline']['value'], 'frac', d['roofline']['frac'], d['clocks'])"
KPS=$(python -c "import json; print(json.load(open('gpurun_out/${TAG}_bench.json'))['kernels_per_step'])")
Q="--no-cpu --no-parity --eager-gpu 0"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s $((KPS * 2)) -c $KPS --csv \
    --log-file gpurun_out/${TAG}_b64_ncu_launches_raw.csv python bench.py --steps 1 --warmup 3 $Q > gpurun_out/${TAG}_b64_ncu.log 2>&1
python tools/ncu_summary.py launches gpurun_out/${TAG}_b64_ncu_launches_raw.csv gpurun_out/${TAG}_b64_launches_summary.csv | head -14
( timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_ops.py tests/test_gpu_tc.py -x -q \
    -k "objective_kernels or conv3d_as_one or token_taps or (fused_layernorm and auto and 389) or (temporal_attn_seq and 24)" 2>&1 | tail -12 ) > gpurun_out/${TAG}_compute_sanitizer_memcheck_new_kernels.txt
tail -6 gpurun_out/${TAG}_compute_sanitizer_memcheck_new_kernels.txt
( MAGE_CUDA_GRAPH=0 timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -x -q \
    -k "forward_loss and tc and forward_L4_b2" 2>&1 | tail -12 ) > gpurun_out/${TAG}_compute_sanitizer_memcheck_forward_loss.txt
tail -6 gpurun_out/${TAG}_compute_sanitizer_memcheck_forward_loss.txt
