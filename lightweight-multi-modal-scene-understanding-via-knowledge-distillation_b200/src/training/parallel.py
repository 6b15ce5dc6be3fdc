"""Data-parallel plumbing: one process per GPU, torch.distributed (NCCL on the GPU
box, gloo in the CPU tests).  Frames are independent through the whole model and
BatchNorm is per-replica as in the reference (no SyncBN), so the only data-path
collective of a step is ONE sum-all-reduce of the flat gradient bucket
(2.1 MB for the weighted student); loss terms and the confusion matrix are
reduced once per epoch, off the critical path.  
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
    """(rank, world_size); (0, 1) when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def allreduce_gradients_(flat_grad: torch.Tensor) -> torch.Tensor:
    """Sum the flat gradient bucket over all ranks in place.  The mean (1/world_size)
    is folded into the optimizer kernel's ``grad_scale`` so no extra pass is spent on it."""
    if world()[1] > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM)
    return flat_grad


def reduce_max(value: float, device="cpu") -> float:
    """max over ranks of a host scalar (step time: the slowest rank defines the step)."""
    if world()[1] == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def reduce_sum(value: float, device="cpu") -> float:
    if world()[1] == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def shard_range(n_items: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous [start, stop) share of ``n_items`` independent frames for ``rank``;
    shares differ by at most one item."""
    base, extra = divmod(n_items, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def frame_seed(rank: int, step: int) -> int:
    """Seed of the synthetic batch a rank generates at a step."""
    return 1000 * rank + step
