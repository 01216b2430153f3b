"""Prompt sharding for N GPUs of one box.

The sampling path has no cross-sample dependency (eval-mode BatchNorm uses running stats,
InstanceNorm is per sample), so the prompt list is cut into contiguous per-rank slices and every
rank runs the whole path on its slice: no collective on the data path.  torch.distributed (NCCL on
the GPU box, gloo in the CPU tests) is used only for the start barrier, the max-over-ranks of the
timings and -- when the caller wants the clips on rank 0 -- a gather of the per-rank outputs.

Results must not depend on the number of GPUs, so everything random is drawn for the GLOBAL batch
(on the CPU generator, like mage_model.py:661) and then sliced with the same bounds.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Tuple

import torch


def shard_bounds(total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of `total` prompts owned by `rank`; the first `total % world`
    ranks take one extra (sizes differ by at most one, an empty slice is legal)."""
    if world <= 0 or not (0 <= rank < world) or total < 0:
        raise ValueError(f"bad shard request: total={total} world={world} rank={rank}")
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(batch: Dict[str, torch.Tensor], world: int, rank: int) -> Dict[str, torch.Tensor]:
    """Slice every tensor of a batch dict ('images', 'text', 'speed', ...) along dim 0."""
    n = next(iter(batch.values())).shape[0]
    lo, hi = shard_bounds(n, world, rank)
    return {k: v[lo:hi] for k, v in batch.items()}


def global_noise(batch: int, res: int = 16, seed: Optional[int] = None) -> torch.Tensor:
    """N(0,1) [B,64,res,res] for the GLOBAL batch on the CPU generator (mage_model.py:661).
    With `seed` every rank draws the identical tensor and slices its rows -- no broadcast needed."""
    if seed is None:
        return torch.randn(batch, 64, res, res)
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return torch.randn(batch, 64, res, res, generator=g)


def noise_for_prompts(seed: int, indices, res: int = 16) -> torch.Tensor:
    """N(0,1) [len(indices), 64, res, res] where row i depends only on (seed, indices[i]) -- the GLOBAL index of the prompt in the
    dataset.  With a seed, a prompt's AdaIN noise (mage_model.py:661) -- and therefore its clip -- is the same whichever rank,
    batch or world size it is generated in (results must not depend on the number of GPUs)."""
    out = torch.empty(len(indices), 64, res, res)
    g = torch.Generator(device="cpu")
    for j, idx in enumerate(indices):
        g.manual_seed((int(seed) * 1000003 + int(idx)) & 0x7FFFFFFFFFFFFFFF)
        out[j] = torch.randn(64, res, res, generator=g)
    return out


def env_world() -> Tuple[int, int, int]:
    """(rank, world, local_rank) as torchrun exports them; (0, 1, 0) for a plain launch."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init_distributed(backend: str, device: Optional[torch.device] = None) -> Tuple[int, int]:
    """Join the torchrun rendezvous when WORLD_SIZE > 1; returns (rank, world)."""
    import torch.distributed as dist

    rank, world, _ = env_world()
    if world > 1 and not dist.is_initialized():
        kw = {"device_id": device} if (backend == "nccl" and device is not None) else {}
        dist.init_process_group(backend, **kw)
    return rank, world


def max_over_ranks(value: float, device: torch.device) -> float:
    import torch.distributed as dist

    t = torch.tensor([value], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device: torch.device) -> float:
    import torch.distributed as dist

    t = torch.tensor([value], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def validation_loss(loss_fn, batches, device: torch.device, distributed: Optional[bool] = None) -> float:
    """The reference's periodic validation (main_mage.py:163-182): the mean over THIS rank's batches of `loss_fn(batch)[0]`
    (MAGE.forward's final loss, eval mode, no gradients), then -- when running distributed -- barrier, all_reduce(SUM) of that
    per-rank mean and division by the world size: the one collective the reference's stage-2 driver issues besides DDP's gradient
    all-reduce.  `loss_fn` returns (0-dim tensor, dict) like MAGE.forward; batches are dicts ('video_id' is dropped, :169-170)."""
    import torch.distributed as dist

    total = torch.zeros((), dtype=torch.float32, device=device)
    count = 0
    with torch.no_grad():
        for batch in batches:
            batch = {k: (v.to(device) if torch.is_tensor(v) else v) for k, v in batch.items() if k != "video_id"}
            loss, _ = loss_fn(batch)
            total += loss.detach().to(device=device, dtype=torch.float32)
            count += 1
    if count == 0:
        raise ValueError("validation needs at least one batch on every rank")
    total /= count
    if distributed is None:
        distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    if distributed:
        dist.barrier()
        dist.all_reduce(total, op=dist.ReduceOp.SUM)
        total /= dist.get_world_size()
    return float(total.item())


def gather_to_rank0(x: torch.Tensor, total: int) -> Optional[torch.Tensor]:
    """Concatenate the per-rank slices (dim 0, sizes per shard_bounds) on rank 0; None elsewhere.
    Ragged slices are padded to the largest one for the all_gather and trimmed afterwards."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return x
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [shard_bounds(total, world, r) for r in range(world)]
    most = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros(most, *x.shape[1:], dtype=x.dtype, device=x.device)
    pad[: x.shape[0]] = x
    parts: List[torch.Tensor] = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    if rank != 0:
        return None
    return torch.cat([p[: hi - lo] for p, (lo, hi) in zip(parts, sizes)], dim=0)
