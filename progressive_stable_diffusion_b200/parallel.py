"""Multi-GPU plumbing: one process per GPU, independent (patient x MES level) units, no collective in the loop.

The reference is single-process; every (patient, level) sample is an independent DDIM trajectory
(inference_pipeline_ip.py:263-308, 377-385), so ranks take contiguous blocks of the flattened unit list and only meet at
a barrier / an optional gather of the finished images.
"""

from __future__ import annotations

import os
from typing import List, Optional

import torch
import torch.distributed as dist


def shard_indices(n: int, rank: int, world_size: int) -> List[int]:
    """Contiguous balanced block of ``range(n)`` for ``rank`` (sizes differ by at most one)."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of {world_size}")
    base, extra = divmod(n, world_size)
    start = rank * base + min(rank, extra)
    return list(range(start, start + base + (1 if rank < extra else 0)))


def init_from_env(backend: Optional[str] = None) -> tuple:
    """(rank, local_rank, world_size) from torchrun's environment; initialises torch.distributed when world_size > 1."""
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        if backend == "nccl":
            torch.cuda.set_device(local)
        # (no device_id=: eager NCCL initialisation pins the calling thread's CPU affinity while it runs, and host worker
        # threads created afterwards inherit the narrowed mask - measured 10x slower CPU baseline on rank 0)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local, world


def shutdown() -> None:
    if dist.is_initialized():
        dist.destroy_process_group()


def barrier() -> None:
    if dist.is_initialized():
        dist.barrier()


def max_over_ranks(value: float, device: Optional[torch.device] = None) -> float:
    if not dist.is_initialized():
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device if dist.get_backend() == "nccl" else None)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device: Optional[torch.device] = None) -> float:
    if not dist.is_initialized():
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device if dist.get_backend() == "nccl" else None)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def gather_to_rank0(x: torch.Tensor) -> Optional[torch.Tensor]:
    """Concatenate equally-shaped per-rank tensors on rank 0 (finished uint8 / fp32 images); None elsewhere."""
    if not dist.is_initialized():
        return x
    world = dist.get_world_size()
    bufs = [torch.empty_like(x) for _ in range(world)] if dist.get_rank() == 0 else None
    dist.gather(x, bufs, dst=0)
    return torch.cat(bufs, dim=0) if bufs is not None else None
