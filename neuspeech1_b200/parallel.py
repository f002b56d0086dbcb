"""Data-parallel plumbing: one process per GPU, batch shards per rank, ONE all-reduce of the flat fp32
trainable-gradient buffer per step.  torch.distributed only carries the collective (NCCL over NVLink on the B200 box, gloo in
the CPU tests); the reference gets the same effect implicitly from HF Trainer + DDP (finetune.py:119-122,248)."""
from __future__ import annotations

import os
from typing import Iterator, List, Optional

import torch
import torch.distributed as dist


class DataParallel:
    def __init__(self, backend: Optional[str] = None, device: Optional[torch.device] = None):
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.owns_group = False
        if self.world > 1 and not dist.is_initialized():
            backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
            kw = {"device_id": device} if (backend == "nccl" and device is not None) else {}
            dist.init_process_group(backend, **kw)
            self.owns_group = True

    def all_reduce_mean(self, flat: torch.Tensor) -> torch.Tensor:
        """In-place mean over ranks of the flat gradient buffer (sum all-reduce, then 1/world)."""
        if self.world > 1:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM)
            flat.mul_(1.0 / self.world)
        return flat

    def shard(self, n_items: int, epoch: int = 0, shuffle: bool = True, seed: int = 0, drop_last: bool = True) -> List[int]:
        """DistributedSampler-equivalent: a seeded permutation, padded/truncated to a multiple of world, strided by rank."""
        g = torch.Generator().manual_seed(seed + epoch)
        idx = torch.randperm(n_items, generator=g).tolist() if shuffle else list(range(n_items))
        if drop_last:
            idx = idx[: (n_items // self.world) * self.world]
        else:
            pad = (-len(idx)) % self.world
            idx = idx + idx[:pad]
        return idx[self.rank::self.world]

    def max_over_ranks(self, value: float, device=None) -> float:
        if self.world == 1:
            return value
        t = torch.tensor([value], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    def barrier(self):
        if self.world > 1:
            dist.barrier()

    def close(self):
        if self.owns_group and dist.is_initialized():
            dist.destroy_process_group()


def linear_warmup_decay(step: int, base_lr: float, warmup_steps: int, total_steps: int) -> float:
    """HF `get_linear_schedule_with_warmup` (lr_scheduler_type='linear', finetune.py:236-237): value for optimizer step `step`."""
    if step < warmup_steps:
        return base_lr * step / max(1, warmup_steps)
    return base_lr * max(0.0, (total_steps - step) / max(1, total_steps - warmup_steps))
