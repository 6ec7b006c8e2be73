"""torch_scatter (2.0.2) API subset used by scatter_v2 — projects/mmdet3d_plugin/ops/sst_ops.py:168,170:
    torch_scatter.scatter_max(feat, unq_inv, dim=0) -> (out, argmax)
    torch_scatter.scatter(feat, unq_inv, dim=0, reduce='mean' | 'sum') -> out
Same names, argument meaning and results: out has index.max()+1 rows (or dim_size), empty segments give 0
and argmax == N, argmax ties resolve to the lowest source row (torch_scatter's sequential CPU rule).

Autograd (first step) follows torch_scatter's own rules: the gradient of a maximum goes to its argmax
row (empty segments, argmax == N, drop theirs), sum broadcasts the segment's gradient to its rows, mean broadcasts it divided by
the clamped row count.  The broadcasts run through the library's row gather; the argmax routing is a torch `scatter_` (backward
only, not on the forward hot path).  argmax is marked non-differentiable."""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from .. import ops

__version__ = "2.0.2+fsfb"


def _prep(src: torch.Tensor, index: torch.Tensor, dim: int, out, dim_size: Optional[int]):
    if out is not None:
        raise NotImplementedError("torch_scatter shim: `out=` accumulation is not used by the FSF path")
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


def _index1d(index: torch.Tensor) -> torch.Tensor:
    return index if index.dim() == 1 else index[(slice(None),) + (0,) * (index.dim() - 1)]


class _SegmentMax(torch.autograd.Function):
    @staticmethod
    def forward(ctx, src2, csr):
        val, arg = ops.segment_reduce(src2, csr, "max", return_argmax=True)
        ctx.save_for_backward(arg)
        ctx.n = src2.size(0)
        ctx.mark_non_differentiable(arg)
        return val, arg

    @staticmethod
    def backward(ctx, grad_val, _grad_arg):
        (arg,) = ctx.saved_tensors
        grad = grad_val.new_zeros((ctx.n + 1, grad_val.size(1)))      # row n collects the empty segments (argmax == N)
        grad.scatter_(0, arg, grad_val.contiguous())
        return grad[: ctx.n], None


class _SegmentSum(torch.autograd.Function):
    @staticmethod
    def forward(ctx, src2, csr, index, mean):
        ctx.save_for_backward(index, csr.offsets)
        ctx.mean = mean
        return ops.segment_reduce(src2, csr, "mean" if mean else "sum")

    @staticmethod
    def backward(ctx, grad_out):
        index, offsets = ctx.saved_tensors
        if ctx.mean:
            cnt = (offsets[1:] - offsets[:-1]).clamp(min=1).to(grad_out.dtype)
            grad_out = grad_out / cnt[:, None]
        return ops.gather_rows(grad_out.contiguous(), index), None, None, None


def scatter_max(src: torch.Tensor, index: torch.Tensor, dim: int = -1, out=None,
                dim_size: Optional[int] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    src2, csr, shape, m = _prep(src, index, dim, out, dim_size)
    if src2.requires_grad and torch.is_grad_enabled():
        val, arg = _SegmentMax.apply(src2, csr)
    else:
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
    if src2.requires_grad and torch.is_grad_enabled():
        res = _SegmentSum.apply(src2, csr, _index1d(index).contiguous(), reduce == "mean")
    else:
        res = ops.segment_reduce(src2, csr, "mean" if reduce == "mean" else "sum")
    return res.reshape((m,) + tuple(shape[1:]))


def scatter_sum(src, index, dim=-1, out=None, dim_size=None):
    return scatter(src, index, dim, out, dim_size, "sum")


scatter_add = scatter_sum


def scatter_mean(src, index, dim=-1, out=None, dim_size=None):
    return scatter(src, index, dim, out, dim_size, "mean")
