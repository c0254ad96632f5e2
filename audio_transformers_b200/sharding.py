"""Multi-GPU plumbing for the front end: the path shards embarrassingly (every clip is independent,
including Whisper's per-clip max), so there is no collective on the data path.  One process per GPU
takes a contiguous shard of the clip list; the only communication is a barrier and a MAX-reduction
of per-rank timings.  Works with any torch.distributed backend (nccl on the
GPU box, gloo in the CPU tests)."""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
    """(rank, world_size); (0, 1) when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n_items: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of ``n_items`` clips for ``rank``: sizes differ by at most one and
    the shards tile [0, n_items) in rank order."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank/world_size {rank}/{world_size}")
    base, rem = divmod(n_items, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def max_over_ranks(value: float, device: torch.device | str = "cpu") -> float:
    """MAX of a per-rank scalar (e.g. device-timed milliseconds) over all ranks."""
    rank, ws = world()
    if ws == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device: torch.device | str = "cpu") -> float:
    rank, ws = world()
    if ws == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def job_throughput(units_this_rank: int, elapsed_ms_this_rank: float, device: torch.device | str = "cpu") -> float:
    """Whole-job units per second: all ranks' units divided by the slowest rank's time."""
    total = sum_over_ranks(units_this_rank, device)
    worst_ms = max_over_ranks(elapsed_ms_this_rank, device)
    return total / (worst_ms * 1e-3)
