"""Frame sharding for multi-GPU runs: one process per GPU, frames are independent units, no data-path
collective (the reference runs samples_per_gpu=1 under torch.distributed.launch,
tools/dist_test.sh:8-11).  The only collective is the MAX over ranks of the timed interval."""
from __future__ import annotations

from typing import List

import torch
import torch.distributed as dist


def frame_ids(rank: int, world: int, n_frames: int) -> List[int]:
    """Frames of a job owned by `rank` (frame i → rank i mod world, as DistributedSampler assigns them)."""
    assert 0 <= rank < world
    return list(range(rank, n_frames, world))


def frame_seed(rank: int, local_index: int, stride: int = 16) -> int:
    """Seed of the `local_index`-th synthetic frame of `rank` (disjoint streams per rank)."""
    return rank * stride + local_index


def max_over_ranks(values: torch.Tensor) -> torch.Tensor:
    """Element-wise MAX over ranks (timings are reported as the slowest rank's)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(values, op=dist.ReduceOp.MAX)
    return values


def throughput(units_per_rank: int, world: int, max_seconds: float) -> float:
    """Whole-job units/s: all ranks' units over the slowest rank's time."""
    return units_per_rank * world / max_seconds
