"""Frame sharding across ranks: inference is "replicas only" -- frames are independent, each
rank owns a disjoint slice of the frame stream, there is no data-path collective; the only communication is the
barrier + max-over-ranks reduction of the step time used for reporting."""
from __future__ import annotations

from typing import List

import torch
import torch.distributed as dist


def frames_for_rank(rank: int, world: int, frames_per_rank: int, step: int = 0) -> List[int]:
    """Frame ids (seeds of the synthetic generator / indices of a dataset) owned by `rank` at `step`:
    a global stream of frames dealt in contiguous blocks of `frames_per_rank`."""
    base = (step * world + rank) * frames_per_rank
    return list(range(base, base + frames_per_rank))


def max_over_ranks(value: float, device="cpu") -> float:
    """MAX all-reduce of a scalar (step time): every multi-GPU number is the slowest rank's."""
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def total_frames(world: int, frames_per_rank: int, steps: int) -> int:
    return world * frames_per_rank * steps
