"""Autograd for the gather-GEMM (sparse convolution / Linear without epilogue) — SURVEY.md section 8f rank 4, second step.

    y = sparse_conv(a, w, nbr)        y[r] = sum_k a[nbr[k][r]] @ w[k].T        (nbr None: plain Linear, w [cout, cin])

* forward and the input gradient run through the library's own persistent tcgen05 kernel: the input gradient of a sparse
  convolution is the same gather-GEMM over the TRANSPOSED rulebook (`inv[k][nbr[k][r]] = r`, well defined because for a fixed
  offset distinct outputs read distinct inputs — true for submanifold, strided and inverse convolutions alike) with `w[k]`
  transposed:  da[j] = sum_k dy[inv[k][j]] @ w[k];
* the weight gradient dw[k] = dy[rows_k].T @ a[nbr[k][rows_k]] is one native launch for all offsets (`fsfb_conv_wgrad`, fp32 CUDA
  cores, deterministic row splits; round 1 ran 27 torch.matmul calls);
* the fused epilogues (folded BatchNorm, LayerNorm, activations, residual) are inference forms and have no backward: training
  composes this function with torch's own norm / activation modules.
Rulebook transposition uses torch indexing (backward only)."""
from __future__ import annotations

from typing import Optional

import torch

from . import ops


def transpose_rulebook(nbr: torch.Tensor, n_in: int) -> torch.Tensor:
    """nbr [koff, n_out] i32 (input row read by offset k of output r, < 0 none) → inv [koff, n_in] i32 (output row that reads
    input j through offset k, -1 none)."""
    inv = torch.full((nbr.size(0), n_in), -1, dtype=torch.int32, device=nbr.device)
    k_idx, r_idx = torch.nonzero((nbr >= 0) & (nbr < n_in), as_tuple=True)
    inv[k_idx, nbr[k_idx, r_idx].long()] = r_idx.to(torch.int32)
    return inv


class _SparseConv(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, w, nbr, order, nbr_ro, symmetric):
        ctx.save_for_backward(a, w, nbr, order, nbr_ro)
        ctx.symmetric = bool(symmetric)
        return ops.gather_gemm(a.contiguous(), ops.gemm_prepack(w.detach()), nbr=nbr, row_order=order, nbr_ro=nbr_ro)

    @staticmethod
    def backward(ctx, grad_out):
        a, w, nbr, order, nbr_ro = ctx.saved_tensors
        w3 = w if w.dim() == 3 else w[None]
        g = grad_out.contiguous()
        grad_a = grad_w = None
        if ctx.needs_input_grad[0]:
            wt = w3.detach().transpose(1, 2).contiguous()                  # [koff, cin, cout]: the "weights" of the transposed conv
            if nbr is not None and ctx.symmetric:
                # submanifold rulebook: offset k of output r reads input j exactly when offset koff-1-k of output j reads input r,
                # so the transposed rulebook is the rulebook itself with the offsets reversed — its row order and row-ordered
                # table are reused and no transposition runs
                grad_a = ops.gather_gemm(g, ops.gemm_prepack(wt.flip(0).contiguous()), nbr=nbr, row_order=order, nbr_ro=nbr_ro)
            else:
                inv = transpose_rulebook(nbr, a.size(0)) if nbr is not None else None
                grad_a = ops.gather_gemm(g, ops.gemm_prepack(wt), nbr=inv)
            if grad_a.size(0) != a.size(0):                                # Linear: rows == a rows by construction
                raise RuntimeError("sparse_conv backward: input gradient has the wrong number of rows")
            grad_a = grad_a[:, : a.size(1)]
        if ctx.needs_input_grad[1]:   # native: csrc/conv_wgrad.cu (pairs compacted per row block, deterministic row splits)
            grad_w = ops.conv_wgrad(a.contiguous(), g, nbr, w3.size(0)).view_as(w)
        return grad_a, grad_w, None, None, None, None


def sparse_conv(a: torch.Tensor, w: torch.Tensor, nbr: Optional[torch.Tensor] = None, order: Optional[torch.Tensor] = None,
                nbr_ro: Optional[torch.Tensor] = None, symmetric: bool = False) -> torch.Tensor:
    """Differentiable gather-GEMM: a [n_in, cin] f32, w [koff, cout, cin] (or [cout, cin] with nbr None), nbr [koff, n_out] i32.
    order / nbr_ro: the rulebook's row order and row-ordered table (modules.Rulebook); symmetric: the rulebook is a submanifold
    one (n_out == n_in, offsets in mirrored order), whose transpose is itself with the offsets reversed."""
    if nbr is None and w.dim() == 3 and w.size(0) != 1:
        raise ValueError("a weight with several offsets needs a rulebook")
    if symmetric and (nbr is None or nbr.size(1) != a.size(0)):
        raise ValueError("a symmetric (submanifold) rulebook has as many outputs as inputs")
    return _SparseConv.apply(a, w, nbr, order, nbr_ro, symmetric)
