"""Column sharding for multi-GPU runs.

Columns are independent in gas optics, so the path shards by contiguous column
ranges with NO data-path collective: one process per GPU, each owning
`shard_columns(...)` of the global column range.  torch.distributed is used only
to bracket timings (barrier + MAX over ranks).
"""
from __future__ import annotations

import os
from typing import Tuple


def shard_columns(ncol: int, world_size: int, rank: int) -> Tuple[int, int]:
    """[c0, c1) of `rank`: ceil-sized contiguous shards, matching the C host path
    (run_host in ecckd_api.cu: per_dev = ceil(ncol / ndev))."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad world_size / rank")
    per = (ncol + world_size - 1) // world_size
    c0 = min(rank * per, ncol)
    return c0, min(c0 + per, ncol)


def dist_env():
    """(rank, local_rank, world_size) from the torchrun environment (defaults: single process)."""
    return (int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)))


def max_over_ranks(value: float, device=None) -> float:
    """MAX of a Python float over all ranks (identity when torch.distributed is not initialised)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device=None) -> float:
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def barrier():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
