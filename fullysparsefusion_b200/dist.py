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


# ---- training-side collectives (host logic only, no backward kernels yet) --------------------
def _world() -> int:
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def coalesced_all_reduce(tensors: List[torch.Tensor], op=None) -> List[torch.Tensor]:
    """ONE all-reduce for a list of same-dtype tensors (flatten → reduce → views copied back in place).  NVSwitch collectives are
    launch-latency bound at these sizes, so k small reductions cost k latencies; one bucket costs one."""
    if not tensors or _world() == 1:
        return tensors
    op = dist.ReduceOp.SUM if op is None else op
    flat = torch.cat([t.reshape(-1) for t in tensors])
    dist.all_reduce(flat, op=op)
    off = 0
    for t in tensors:
        t.copy_(flat[off:off + t.numel()].view_as(t))
        off += t.numel()
    return tensors


def reduce_means(values: List[torch.Tensor]) -> List[torch.Tensor]:
    """mmdet.core.reduce_mean for several scalars at once: the heads call it once per loss normaliser (sparse_cluster_head.py:142,160;
    sparse_cluster_head_v2.py:235,251; frustum_cluster_head.py:184,203 — six or more single-float all-reduces per step upstream).
    Returns new tensors = mean over ranks, inputs untouched, exactly as reduce_mean does for each."""
    w = _world()
    if w == 1 or not values:
        return [v.clone() for v in values]
    flat = torch.stack([v.detach().reshape(()).to(torch.float32) for v in values]) / w
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    return [flat[i].to(values[i].dtype) for i in range(len(values))]


class _AllReduceSum(torch.autograd.Function):
    """Differentiable sum all-reduce: the gradient of a sum over ranks is the sum over ranks of the gradients (detectron2's
    differentiable AllReduce, which NaiveSyncBatchNorm relies on)."""

    @staticmethod
    def forward(ctx, x):
        y = x.clone()
        dist.all_reduce(y, op=dist.ReduceOp.SUM)
        return y

    @staticmethod
    def backward(ctx, g):
        g = g.clone()
        dist.all_reduce(g, op=dist.ReduceOp.SUM)
        return g


def naive_sync_bn_stats(x: torch.Tensor):
    """Batch statistics of `naiveSyncBN1d` in training mode [UPSTREAM-RECALL: the fork's class is detectron2's NaiveSyncBatchNorm for
    1-d inputs]: per-rank mean and mean of squares concatenated into ONE [2C] vector, all-reduced, divided by the world size (ranks
    weigh equally whatever their row counts), var = E[x^2] - E[x]^2.  Layers are sequentially dependent, so the [2C] vector per
    layer is the coalescing limit in the forward pass.  Returns (mean [C], var [C])."""
    assert x.dim() == 2 and x.size(0) > 0
    c = x.size(1)
    vec = torch.cat([x.mean(0), (x * x).mean(0)])
    w = _world()
    if w > 1:
        vec = _AllReduceSum.apply(vec) / w   # differentiable: the statistics are part of the training graph
    mean, meansq = vec[:c], vec[c:]
    return mean, meansq - mean * mean
