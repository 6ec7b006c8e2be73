"""Experimental plane-split execution of the 3x3x3 sparse convolutions (modules.PLANE_SPLIT): three gather-GEMM passes over the
z-planes of the rulebook, chained through the residual input, equal the one-pass convolution.

CPU: the host logic (rulebook / packed-weight slicing, scale folding, residual chaining) runs with `ops.gemm_prepack /
gather_gemm / rulebook_row_order` replaced by numpy-oracle stand-ins (test-only).  GPU: the real kernels, gated."""
import numpy as np
import pytest
import torch

from fullysparsefusion_b200 import modules as M
from fullysparsefusion_b200 import ops, synth
from oracle import fsf_oracle as O


def _scene(n_pts, seed):
    pts = synth.ring_points(n_pts, sweeps=2, seed=seed)
    uniq, _, _ = O.unique_rows(O.voxelize(pts, synth.NUSC_VOXEL, synth.NUSC_RANGE, floor_mode=1))
    c4 = np.concatenate([np.zeros((len(uniq), 1), np.int64), uniq], 1)
    nbr = O.conv_rulebook(c4, c4, (1, 40, 512, 512), (3, 3, 3), (1, 1, 1), (1, 1, 1))
    return c4, nbr.astype(np.int32)


def _module(cin, cout, act=True):
    torch.manual_seed(cin + cout)
    m = M.SparseConvModule(cin, cout, dict(type="BN1d", eps=1e-3, momentum=0.01), act=act)
    with torch.no_grad():
        m.bn.running_mean.normal_(0, 0.3)
        m.bn.running_var.uniform_(0.5, 2.0)
        m.bn.weight.uniform_(0.5, 1.5)
        m.bn.bias.normal_(0, 0.2)
    return m.eval()


def _compare(device, monkeypatch, rtol, atol):
    c4, nbr_np = _scene(4000, 3)
    rows = nbr_np.shape[1]
    assert (nbr_np[:9] >= 0).any() and (nbr_np[18:] >= 0).any()
    nbr = torch.from_numpy(nbr_np).to(device)
    monkeypatch.setattr(M, "PLANE_SPLIT_MIN_ROWS", 1)
    for cin, cout, act, with_res in [(32, 64, True, False), (16, 16, True, True), (48, 128, False, False)]:
        mod = _module(cin, cout, act).to(device)
        rng = np.random.default_rng(cin)
        x = torch.from_numpy(np.maximum(rng.standard_normal((rows, cin)), 0).astype(np.float32)).to(device)
        res = torch.from_numpy(rng.standard_normal((rows, cout)).astype(np.float32)).to(device) if with_res else None
        monkeypatch.setattr(M, "PLANE_SPLIT", False)
        want = mod(x, M.Rulebook(nbr, sort_rows=False), residual=res)
        monkeypatch.setattr(M, "PLANE_SPLIT", True)
        mod.refresh()
        got = mod(x, M.Rulebook(nbr, sort_rows=False), residual=res)
        torch.testing.assert_close(got, want, rtol=rtol, atol=atol)
        s = mod.bn.weight / torch.sqrt(mod.bn.running_var + mod.bn.eps)
        direct = O.gather_gemm(x.cpu().numpy(), mod.weight.detach().cpu().numpy(), nbr_np, norm="affine", norm_w=s.detach().cpu().numpy(),
                               norm_b=(mod.bn.bias - mod.bn.running_mean * s).detach().cpu().numpy(),
                               residual=None if res is None else res.cpu().numpy(), act="relu" if act else None)
        np.testing.assert_allclose(got.cpu().numpy(), direct, rtol=rtol, atol=atol)
    # a residual added after the activation is not expressible as a chain: that call keeps the one-pass form
    calls = []
    real = ops.gather_gemm
    monkeypatch.setattr(ops, "gather_gemm", lambda *a, **k: (calls.append(k.get("nbr").size(0)), real(*a, **k))[1])
    mod(x, M.Rulebook(nbr, sort_rows=False), residual=torch.zeros(rows, 128, device=device), residual_post=True)
    assert calls == [27]
    calls.clear()
    mod(x, M.Rulebook(nbr, sort_rows=False))
    assert calls == [9, 9, 9]


def test_plane_split_host_logic_on_cpu(monkeypatch):
    def prepack(w, keep_raw=False):
        w3 = (w if w.dim() == 3 else w[None]).contiguous()
        koff, cout, cin = w3.shape
        size = koff * ((cin + 31) // 32) * 2 * ((cout + 15) // 16 * 16) * 128          # GemmShape::total_bytes for one column tile
        return ops.PackedWeight(torch.zeros(size, dtype=torch.uint8), koff, cin, cout, w3)

    def gather_gemm(a, w, nbr=None, norm=None, norm_w=None, norm_b=None, residual=None, act=None, out=None, row_order=None, **kw):
        assert w.data.numel() == w.koff * ((w.cin + 31) // 32) * 2 * ((w.cout + 15) // 16 * 16) * 128 and w.raw.size(0) == w.koff
        assert nbr.size(0) == w.koff and (row_order is None or sorted(row_order.tolist()) == list(range(nbr.size(1))))
        y = O.gather_gemm(a.numpy(), w.raw.numpy(), nbr.numpy(), norm=norm, norm_w=None if norm_w is None else norm_w.numpy(),
                          norm_b=None if norm_b is None else norm_b.numpy(), residual=None if residual is None else residual.numpy(), act=act)
        y = torch.from_numpy(y.astype(np.float32))
        if out is not None:
            out.copy_(y)
            return out
        return y

    def row_order(nbr):
        m = sum((nbr[k] >= 0).to(torch.int64) << k for k in range(nbr.size(0)))
        return torch.argsort(m, stable=True).to(torch.int32)

    monkeypatch.setattr(ops, "gemm_prepack", prepack)
    monkeypatch.setattr(ops, "gather_gemm", gather_gemm)
    monkeypatch.setattr(ops, "rulebook_row_order", row_order)
    _compare("cpu", monkeypatch, 1e-5, 1e-5)


@pytest.mark.gpu
def test_plane_split_on_device(cuda, monkeypatch):
    _compare("cuda:0", monkeypatch, 1e-4, 2e-5)
