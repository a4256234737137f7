"""Stream sharding across GPUs.

Frames of one stream are strictly recurrent (scripts/training/models.py:752-764,
823), so the only multi-GPU mode is data parallel over independent streams:
stream s runs on rank (s mod world) for its whole life, its recurrent state
never leaves that GPU, and there is no collective on the frame path.
torch.distributed is used only to fence the timed region and to take the
max-over-ranks time (NCCL on GPUs, gloo in the CPU tests).
"""

from __future__ import annotations

from typing import List


def streams_for_rank(total_streams: int, world: int, rank: int) -> List[int]:
    """Round-robin assignment: global stream ids owned by `rank`."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank / world")
    return list(range(rank, total_streams, world))


def max_over_ranks(value: float, dist=None, device=None) -> float:
    """All-reduce MAX of a scalar (identity when not distributed)."""
    if dist is None or not dist.is_initialized():
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def aggregate_fps(streams_per_rank: int, world: int, steps: int, max_seconds: float) -> float:
    """Whole-job throughput: frames produced by all ranks / slowest rank's time."""
    return streams_per_rank * world * steps / max_seconds
