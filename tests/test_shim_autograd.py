"""Autograd of the torch_scatter shim (first step) against torch's own autograd of the same reductions.

The backward glue is ordinary torch code around two library ops, so the CPU test runs it with `ops.build_csr / segment_reduce /
gather_rows` replaced by numpy-oracle stand-ins (test-only; the product ops refuse CPU tensors); the GPU test runs the real thing."""
import numpy as np
import pytest
import torch

from fullysparsefusion_b200 import ops
from fullysparsefusion_b200.shims import torch_scatter as TS
from oracle import fsf_oracle as O


def _reference(src, index, m, mode):
    n, c = src.shape
    if mode == "max":
        out = torch.full((m, c), float("-inf"), dtype=src.dtype, device=src.device)
        out = out.scatter_reduce(0, index[:, None].expand(n, c), src, "amax", include_self=True)
        return torch.where(torch.isinf(out), torch.zeros_like(out), out)      # empty segments → 0 (no gradient)
    out = torch.zeros((m, c), dtype=src.dtype, device=src.device).index_add(0, index, src)
    if mode == "mean":
        cnt = torch.bincount(index, minlength=m).clamp(min=1).to(src.dtype)
        out = out / cnt[:, None]
    return out


def _check(device):
    g = torch.Generator().manual_seed(3)
    cases = [(500, 7, 40, 40), (300, 33, 64, 50), (64, 4, 5, 5)]                                # m > hi: trailing empty segments
    if device == "cpu":
        cases.append((1, 3, 4, 2))
    for n, c, m, hi in cases:
        src = torch.randn(n, c, generator=g).to(device)
        index = torch.randint(0, hi, (n,), generator=g).to(device)
        weight = torch.randn(m, c, generator=g).to(device)
        for mode in ("max", "mean", "sum"):
            a = src.clone().requires_grad_(True)
            b = src.clone().requires_grad_(True)
            if mode == "max":
                got, arg = TS.scatter_max(a, index, dim=0, dim_size=m)
                assert not arg.requires_grad and arg.dtype == torch.int64
            else:
                got = TS.scatter(a, index, dim=0, dim_size=m, reduce=mode)
            want = _reference(b, index, m, mode)
            torch.testing.assert_close(got, want, rtol=1e-5, atol=1e-5)
            (got * weight).sum().backward()
            (want * weight).sum().backward()
            assert a.grad.shape == src.shape
            torch.testing.assert_close(a.grad, b.grad, rtol=1e-5, atol=1e-6)
    # no graph when nothing requires grad / under no_grad
    out = TS.scatter(src, index, dim=0, reduce="mean")
    assert not out.requires_grad
    with torch.no_grad():
        assert not TS.scatter_max(src.clone().requires_grad_(True), index, dim=0)[0].requires_grad
    # a 3-d source keeps its trailing shape through forward and backward
    x = torch.randn(50, 3, 4, generator=g).to(device).requires_grad_(True)
    idx = torch.randint(0, 6, (50,), generator=g).to(device)
    y, _ = TS.scatter_max(x, idx, dim=0)
    y.sum().backward()
    assert y.shape == (int(idx.max()) + 1, 3, 4) and x.grad.shape == x.shape and float(x.grad.sum()) == y.numel()


def test_backward_glue_on_cpu(monkeypatch):
    def build_csr(index, m):
        index = index.to(torch.int64)
        perm = torch.argsort(index, stable=True)
        offsets = torch.zeros(m + 1, dtype=torch.int32)
        offsets[1:] = torch.cumsum(torch.bincount(index, minlength=m), 0)
        return ops.SegmentCSR(offsets, perm.to(torch.int32), index[perm].to(torch.int32), index.numel(), m)

    def segment_reduce(feat, csr, mode, return_argmax=False):
        index = np.empty(csr.n, np.int64)
        index[csr.perm.numpy()] = csr.seg.numpy()
        f = feat.detach().numpy()
        if mode == "max":
            val, arg = O.scatter_max(f, index, m=csr.m)
            return (torch.from_numpy(val), torch.from_numpy(arg)) if return_argmax else torch.from_numpy(val)
        fn = O.scatter_mean if mode == "mean" else O.scatter_sum
        return torch.from_numpy(fn(f, index, m=csr.m))

    monkeypatch.setattr(ops, "build_csr", build_csr)
    monkeypatch.setattr(ops, "segment_reduce", segment_reduce)
    monkeypatch.setattr(ops, "gather_rows", lambda src, idx, fill=0.0, out=None: src[idx.long()])
    _check("cpu")


@pytest.mark.gpu
def test_backward_on_device(cuda):
    _check("cuda:0")
