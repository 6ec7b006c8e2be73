"""torch_scatter (2.0.2) API subset used by scatter_v2 — projects/mmdet3d_plugin/ops/sst_ops.py:168,170:
    torch_scatter.scatter_max(feat, unq_inv, dim=0) -> (out, argmax)
    torch_scatter.scatter(feat, unq_inv, dim=0, reduce='mean' | 'sum') -> out
Same names, argument meaning and results: out has index.max()+1 rows (or dim_size), empty segments give 0
and argmax == N, argmax ties resolve to the lowest source row (torch_scatter's sequential CPU rule)."""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from .. import ops

__version__ = "2.0.2+fsfb"


def _prep(src: torch.Tensor, index: torch.Tensor, dim: int, out, dim_size: Optional[int]):
    if out is not None:
        raise NotImplementedError("torch_scatter shim: `out=` accumulation is not used by the FSF path")
    if src.requires_grad:
        raise NotImplementedError("torch_scatter shim: forward/inference only in this round")
    if dim < 0:
        dim += src.dim()
    if dim != 0:
        raise NotImplementedError("torch_scatter shim: the FSF path scatters over dim 0 only (sst_ops.py:168,170)")
    if index.dim() != 1:
        if index.dim() == src.dim() and index.size(0) == src.size(0):
            index = index[(slice(None),) + (0,) * (index.dim() - 1)]  # broadcast index: same id along a row
        else:
            raise ValueError("torch_scatter shim: index must be 1-D over dim 0")
    assert index.numel() == src.size(0), (index.shape, src.shape)
    shape = src.shape
    src2 = src.reshape(shape[0], -1)
    if src2.dtype != torch.float32:
        raise NotImplementedError("torch_scatter shim: float32 features only")
    if dim_size is None:
        dim_size = int(index.max().item()) + 1 if index.numel() else 0  # the sync torch_scatter also pays
    csr = ops.build_csr(index, dim_size)
    return src2, csr, shape, dim_size


def scatter_max(src: torch.Tensor, index: torch.Tensor, dim: int = -1, out=None,
                dim_size: Optional[int] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    src2, csr, shape, m = _prep(src, index, dim, out, dim_size)
    val, arg = ops.segment_reduce(src2, csr, "max", return_argmax=True)
    return val.reshape((m,) + tuple(shape[1:])), arg.reshape((m,) + tuple(shape[1:]))


def scatter_min(src, index, dim=-1, out=None, dim_size=None):
    v, a = scatter_max(-src, index, dim, out, dim_size)
    return -v, a


def scatter(src: torch.Tensor, index: torch.Tensor, dim: int = -1, out=None, dim_size: Optional[int] = None,
            reduce: str = "sum") -> torch.Tensor:
    if reduce == "max":
        return scatter_max(src, index, dim, out, dim_size)[0]
    if reduce == "min":
        return scatter_min(src, index, dim, out, dim_size)[0]
    if reduce not in ("sum", "add", "mean"):
        raise ValueError(f"torch_scatter shim: unsupported reduce {reduce!r}")
    src2, csr, shape, m = _prep(src, index, dim, out, dim_size)
    res = ops.segment_reduce(src2, csr, "mean" if reduce == "mean" else "sum")
    return res.reshape((m,) + tuple(shape[1:]))


def scatter_sum(src, index, dim=-1, out=None, dim_size=None):
    return scatter(src, index, dim, out, dim_size, "sum")


scatter_add = scatter_sum


def scatter_mean(src, index, dim=-1, out=None, dim_size=None):
    return scatter(src, index, dim, out, dim_size, "mean")
