"""Autograd of the gather-GEMM (fullysparsefusion_b200/autograd.py) against torch's own autograd of the same sum.

The CPU test runs the backward glue (rulebook transposition, transposed weights, per-offset weight gradients) with
`ops.gemm_prepack / gather_gemm` replaced by a torch stand-in (test-only; the product ops refuse CPU tensors); the GPU test runs
the real kernels (gated: written after the GPU budget was spent)."""
import types

import numpy as np
import pytest
import torch

from fullysparsefusion_b200 import autograd as AG
from fullysparsefusion_b200 import ops


def _reference(a, w, nbr):
    w3 = w if w.dim() == 3 else w[None]
    if nbr is None:
        return a @ w3[0].t()
    out = 0
    for k in range(w3.size(0)):
        src = nbr[k].long()
        ok = (src >= 0) & (src < a.size(0))
        rows = a[src.clamp(0, a.size(0) - 1)] * ok[:, None]
        out = out + rows @ w3[k].t()
    return out


def _cases(device):
    g = torch.Generator().manual_seed(5)
    out = []
    # submanifold-like: n_out == n_in, symmetric-ish random rulebook built from an injective map per offset
    for n_in, n_out, cin, cout, koff in [(200, 200, 32, 64, 27), (150, 90, 64, 32, 8), (64, 300, 16, 16, 8)]:
        nbr = torch.full((koff, n_out), -1, dtype=torch.int32)
        for k in range(koff):
            m = min(n_in, n_out)
            src = torch.randperm(n_in, generator=g)[:m]                  # injective per offset, as real rulebooks are
            dst = torch.randperm(n_out, generator=g)[:m]
            keep = torch.rand(m, generator=g) < 0.4
            nbr[k, dst[keep]] = src[keep].to(torch.int32)
        a = torch.randn(n_in, cin, generator=g)
        w = torch.randn(koff, cout, cin, generator=g) / (cin * koff) ** 0.5
        out.append((a.to(device), w.to(device), nbr.to(device)))
    a = torch.randn(300, 96, generator=g)
    w = torch.randn(64, 96, generator=g) / 96 ** 0.5
    out.append((a.to(device), w.to(device), None))                       # plain Linear
    return out


def _check(device, rtol, atol):
    for a, w, nbr in _cases(device):
        a1, w1 = a.clone().requires_grad_(True), w.clone().requires_grad_(True)
        a2, w2 = a.clone().requires_grad_(True), w.clone().requires_grad_(True)
        y = AG.sparse_conv(a1, w1, nbr)
        want = _reference(a2, w2, nbr)
        torch.testing.assert_close(y, want, rtol=rtol, atol=atol)
        coef = torch.randn(want.shape, generator=torch.Generator().manual_seed(1)).to(device)
        (y * coef).sum().backward()
        (want * coef).sum().backward()
        torch.testing.assert_close(a1.grad, a2.grad, rtol=rtol, atol=atol)
        torch.testing.assert_close(w1.grad, w2.grad, rtol=rtol, atol=atol)
        if nbr is not None:                                              # transposing twice gives the rulebook back
            inv = AG.transpose_rulebook(nbr, a.size(0))
            assert torch.equal(AG.transpose_rulebook(inv, nbr.size(1)), nbr)
    # input gradient only / weight gradient only
    a, w, nbr = _cases(device)[0]
    a1 = a.clone().requires_grad_(True)
    AG.sparse_conv(a1, w, nbr).sum().backward()
    assert a1.grad is not None
    w1 = w.clone().requires_grad_(True)
    AG.sparse_conv(a, w1, nbr).sum().backward()
    assert w1.grad is not None and w1.grad.shape == w.shape


def _wgrad_reference(a, dy, nbr, koff, dtype=torch.float64):
    """dw[k] = dy[rows_k].T @ a[nbr[k][rows_k]] (the 27 matmuls the native kernel replaced), float64 unless told otherwise."""
    a64, g64 = a.to(dtype), dy.to(dtype)
    dw = torch.zeros(koff, dy.size(1), a.size(1), dtype=dtype, device=a.device)
    for k in range(koff):
        if nbr is None:
            dw[k] = g64.t() @ a64
        else:
            rows = torch.nonzero((nbr[k] >= 0) & (nbr[k] < a.size(0)))[:, 0]
            if rows.numel():
                dw[k] = g64[rows].t() @ a64[nbr[k][rows].long()]
    return dw


def test_backward_glue_on_cpu(monkeypatch):
    monkeypatch.setattr(ops, "gemm_prepack", lambda w, keep_raw=False: types.SimpleNamespace(raw=w if w.dim() == 3 else w[None]))
    monkeypatch.setattr(ops, "gather_gemm", lambda a, pw, nbr=None, **kw: _reference(a, pw.raw, nbr))
    monkeypatch.setattr(ops, "conv_wgrad", lambda a, dy, nbr, koff: _wgrad_reference(a, dy, nbr, koff, torch.float32))
    _check("cpu", 1e-5, 1e-5)


@pytest.mark.gpu
def test_backward_on_device(cuda):
    _check("cuda:0", 1e-4, 2e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("n_in,n_out,cin,cout,koff,density", [
    (5000, 5000, 128, 128, 27, 0.2),     # a submanifold level: several row splits
    (40000, 40000, 64, 128, 27, 0.2),
    (3000, 900, 128, 256, 27, 0.3),      # strided: fewer outputs than inputs, four cout tiles
    (900, 3000, 256, 128, 8, 0.5),       # inverse
    (700, 700, 33, 11, 27, 0.3),         # odd widths, partial tiles
    (10000, 10000, 131, 128, 1, 1.0),    # Linear through a table
    (20000, 20000, 133, 64, 0, 1.0),     # Linear, no table
    (1, 1, 16, 16, 27, 1.0),
])
def test_conv_wgrad_kernel(cuda, n_in, n_out, cin, cout, koff, density):
    """csrc/conv_wgrad.cu against per-offset float64 matmuls; twice the same bits (deterministic: no atomics)."""
    g = torch.Generator().manual_seed(n_in + cin + koff)
    a = torch.randn(n_in, cin, generator=g).to(cuda)
    dy = torch.randn(n_out, cout, generator=g).to(cuda)
    nbr = None
    if koff:
        nbr = torch.randint(0, n_in, (koff, n_out), generator=g, dtype=torch.int32)
        nbr[torch.rand(koff, n_out, generator=g) >= density] = -1
        nbr = nbr.to(cuda)
    k = max(koff, 1)
    got = ops.conv_wgrad(a, dy, nbr, k)
    want = _wgrad_reference(a, dy, nbr, k)
    scale = want.abs().max().clamp(min=1e-6)
    assert float((got.double() - want).abs().max() / scale) < 2e-5
    assert torch.equal(got, ops.conv_wgrad(a, dy, nbr, k))
    # strided operands (a row slice of a wider buffer)
    wide = torch.zeros(n_in, cin + 7, device=cuda)
    wide[:, :cin] = a
    assert torch.equal(got, ops.conv_wgrad(wide[:, :cin], dy, nbr, k))
