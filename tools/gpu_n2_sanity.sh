#!/bin/bash
# 2-GPU sanity of the last state: default strong-scaling line and the validation-loss workload under torchrun
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 5 --warmup 3 \
  > gpurun_out/r02ag_bench_n2.json 2> gpurun_out/r02ag_bench_n2.err
python -c "
import json; d=json.loads(open('gpurun_out/r02ag_bench_n2.json').read().strip().splitlines()[-1]); print('n2', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'parity', d['parity']['token_mismatches'], d.get('weak_scaling'))"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --workload c4val --steps 5 --warmup 3 \
  > gpurun_out/r02ag_bench_c4val_n2.json 2> gpurun_out/r02ag_bench_c4val_n2.err
python -c "
import json; d=json.loads(open('gpurun_out/r02ag_bench_c4val_n2.json').read().strip().splitlines()[-1]); print('c4val n2', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'test_loss', d['test_loss'], d['loss_dict'])"
tail -3 gpurun_out/r02ag_bench_c4val_n2.err | cut -c1-300
