#!/bin/bash
# `main_mage.py --split val` (validation loss of a checkpoint, the forward half of the stage-2 driver): 1 process vs torchrun x N over NCCL
mkdir -p gpurun_out /tmp/valck
python - <<'PY'
import torch, yaml
from mage_b200 import synthetic as syn
params = syn.model_params("caterv2", frames_length=16)
open("/tmp/valck/config.yaml", "w").write(yaml.safe_dump({"model": {"target": "modules.mage_model.MAGE", "params": params},
                                                             "data": {"target": "dataload.CATER", "params": {}}}))
torch.save({"state_dict": syn.make_mage_state_dict(params, posterior=True)}, "/tmp/valck/model_best.pth")
PY
N=${1:-2}
{
echo "# 1 process"
python main_mage.py --split val --test_model /tmp/valck/model_best.pth --synthetic 32 --batch-size 8 --seed 9 2>&1 | grep -v Warning | tail -3
echo "# torchrun x $N (NCCL all_reduce of the per-rank mean, main_mage.py:177-180)"
NCCL_DEBUG=INFO NCCL_DEBUG_FILE=gpurun_out/val_nccl_%h_%p.log python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
  main_mage.py --split val --test_model /tmp/valck/model_best.pth --synthetic 32 --batch-size 8 --seed 9 2>&1 | grep -v Warning | tail -3
grep -h "AllReduce\|NVLS\|via P2P\|Connected all" gpurun_out/val_nccl_*.log | head -8
rm -f gpurun_out/val_nccl_*.log
} 2>&1 | tee gpurun_out/val_n${N}.txt
