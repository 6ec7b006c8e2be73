"""GPU parity of the gather-GEMM (tcgen05 3xTF32) and its fused epilogue against the CPU oracle,
the reference-generated build_mlp goldens, and the CUDA-core fp32 cross-check.  Tolerance: 1e-4
relative (north_star) — the 3xTF32 split lands near 1e-6."""
import numpy as np
import pytest
import torch

from fullysparsefusion_b200 import ops
from oracle import fsf_oracle as O
from tests.conftest import load_golden

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-4, 2e-5


def T(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def _rel_err(got, want):
    scale = np.maximum(np.abs(want), 1e-2 * np.abs(want).max() + 1e-12)
    return float(np.max(np.abs(got - want) / scale))


@pytest.mark.parametrize("rows,cin,cout", [(1, 8, 16), (128, 32, 128), (300, 5, 64), (1000, 131, 128), (257, 64, 11),
                                           (513, 180, 128), (200, 768, 1024), (129, 133, 256), (64, 3, 16),
                                           (5000, 128, 33), (333, 10, 131), (100, 896, 1024), (77, 1024, 3)])
def test_linear_plain(cuda, rows, cin, cout):
    rng = np.random.default_rng(rows + cin + cout)
    a = rng.standard_normal((rows, cin)).astype(np.float32)
    w = (rng.standard_normal((cout, cin)) / np.sqrt(cin)).astype(np.float32)
    b = rng.standard_normal(cout).astype(np.float32)
    pw = ops.gemm_prepack(T(w, cuda), keep_raw=True)
    want = O.gather_gemm(a, w, bias=b)
    got = ops.gather_gemm(T(a, cuda), pw, bias=T(b, cuda)).cpu().numpy()
    simt = ops.gather_gemm(T(a, cuda), pw, bias=T(b, cuda), simt=True).cpu().numpy()
    np.testing.assert_allclose(simt, want, rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(got, want, rtol=RTOL, atol=ATOL)
    # 3xTF32 with split accumulators: error relative to the output scale stays ~1e-5 even at K=1024
    # (plain TF32 would sit at ~5e-4)
    assert np.abs(got - want).max() / np.abs(want).max() < 2e-5, np.abs(got - want).max() / np.abs(want).max()


@pytest.mark.parametrize("norm,act", [("ln", "gelu"), ("affine", "relu"), (None, "gelu"), ("ln", None)])
@pytest.mark.parametrize("rows,cin,cout", [(700, 48, 64), (300, 131, 128), (1000, 64, 256), (90, 16, 32)])
def test_linear_epilogue(cuda, norm, act, rows, cin, cout):
    rng = np.random.default_rng(cin * cout)
    a = rng.standard_normal((rows, cin)).astype(np.float32)
    w = (rng.standard_normal((cout, cin)) / np.sqrt(cin)).astype(np.float32)
    b = rng.standard_normal(cout).astype(np.float32)
    nw = rng.uniform(0.5, 1.5, cout).astype(np.float32)
    nb = rng.standard_normal(cout).astype(np.float32)
    res = rng.standard_normal((rows, cout)).astype(np.float32)
    want = O.gather_gemm(a, w, bias=b, norm=norm, norm_w=nw, norm_b=nb, eps=1e-3, residual=res, act=act)
    pw = ops.gemm_prepack(T(w, cuda), keep_raw=True)
    kw = dict(bias=T(b, cuda), norm=norm, norm_w=T(nw, cuda) if norm else None, norm_b=T(nb, cuda) if norm else None,
              eps=1e-3, residual=T(res, cuda), act=act)
    got = ops.gather_gemm(T(a, cuda), pw, **kw).cpu().numpy()
    np.testing.assert_allclose(got, want, rtol=RTOL, atol=ATOL)
    simt = ops.gather_gemm(T(a, cuda), pw, simt=True, **kw).cpu().numpy()
    np.testing.assert_allclose(simt, want, rtol=RTOL, atol=ATOL)
    # standalone row epilogue on the pre-norm sums
    pre = ops.gather_gemm(T(a, cuda), pw)
    kw2 = dict(kw)
    got2 = ops.rownorm_act(pre, **kw2).cpu().numpy()
    np.testing.assert_allclose(got2, want, rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("tag,norm,act", [("mlp_ln_gelu", "ln", "gelu"), ("mlp_head", "ln", "gelu"), ("mlp_bn_relu", "bn", "relu")])
def test_build_mlp_golden(cuda, tag, norm, act):
    """The reference's own build_mlp (sst_ops.py:808-833) outputs, layer by layer through the C-ABI."""
    g = load_golden(tag)
    sd = {k.replace("__", "."): v for k, v in g.items() if k not in ("x", "y")}
    x = T(g["x"], cuda)
    i = 0
    while True:
        if f"{i}.0.weight" in sd:
            pw = ops.gemm_prepack(T(sd[f"{i}.0.weight"], cuda))
            if norm == "ln":
                x = ops.gather_gemm(x, pw, norm="ln", norm_w=T(sd[f"{i}.1.weight"], cuda), norm_b=T(sd[f"{i}.1.bias"], cuda),
                                    eps=1e-3, act=act)
            else:  # eval-mode BN folded to an affine
                scale = sd[f"{i}.1.weight"] / np.sqrt(sd[f"{i}.1.running_var"] + np.float32(1e-3))
                shift = sd[f"{i}.1.bias"] - sd[f"{i}.1.running_mean"] * scale
                x = ops.gather_gemm(x, pw, norm="affine", norm_w=T(scale.astype(np.float32), cuda),
                                    norm_b=T(shift.astype(np.float32), cuda), act=act)
        elif f"{i}.weight" in sd:
            x = ops.gather_gemm(x, ops.gemm_prepack(T(sd[f"{i}.weight"], cuda)), bias=T(sd[f"{i}.bias"], cuda))
        else:
            break
        i += 1
    np.testing.assert_allclose(x.cpu().numpy(), g["y"], rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("rows,a_rows,koff,cin,cout,density", [(500, 400, 27, 16, 32, 0.4), (1000, 1000, 27, 64, 64, 0.3),
                                                                (130, 90, 8, 128, 256, 0.5), (2000, 2500, 27, 128, 128, 0.05),
                                                                (256, 256, 27, 5, 16, 1.0)])
def test_gather_gemm_oracle(cuda, rows, a_rows, koff, cin, cout, density):
    rng = np.random.default_rng(rows + koff)
    a = rng.standard_normal((a_rows, cin)).astype(np.float32)
    w = (rng.standard_normal((koff, cout, cin)) / np.sqrt(cin * koff * density)).astype(np.float32)
    nbr = rng.integers(0, a_rows, (koff, rows)).astype(np.int32)
    nbr[rng.random((koff, rows)) > density] = -1
    if density < 0.1:
        nbr[3] = -1          # a whole offset with no pairs (skipped by the kernel)
        nbr[:, 128:256] = -1  # a whole tile with no input at all → zeros + epilogue
    scale = rng.uniform(0.5, 1.5, cout).astype(np.float32)
    shift = rng.standard_normal(cout).astype(np.float32)
    want = O.gather_gemm(a, w, nbr, norm="affine", norm_w=scale, norm_b=shift, act="relu")
    pw = ops.gemm_prepack(T(w, cuda), keep_raw=True)
    kw = dict(nbr=T(nbr, cuda), norm="affine", norm_w=T(scale, cuda), norm_b=T(shift, cuda), act="relu")
    got = ops.gather_gemm(T(a, cuda), pw, **kw).cpu().numpy()
    np.testing.assert_allclose(got, want, rtol=RTOL, atol=ATOL)
    simt = ops.gather_gemm(T(a, cuda), pw, simt=True, **kw).cpu().numpy()
    np.testing.assert_allclose(simt, want, rtol=RTOL, atol=ATOL)


def test_gather_gemm_large_vs_simt(cuda):
    """Full-size layer (160k voxels x 27 offsets x 128→128) against the fp32 CUDA-core path."""
    g = torch.Generator(device=cuda).manual_seed(0)
    m, koff, c = 160_000, 27, 128
    a = torch.randn(m, c, device=cuda, generator=g)
    w = torch.randn(koff, c, c, device=cuda, generator=g) / (c * 10) ** 0.5
    nbr = torch.randint(0, m, (koff, m), device=cuda, generator=g, dtype=torch.int32)
    nbr[torch.rand(koff, m, device=cuda, generator=g) > 0.37] = -1
    nbr[13] = torch.arange(m, device=cuda, dtype=torch.int32)
    pw = ops.gemm_prepack(w, keep_raw=True)
    got = ops.gather_gemm(a, pw, nbr=nbr, act="relu")
    want = ops.gather_gemm(a, pw, nbr=nbr, act="relu", simt=True)
    torch.testing.assert_close(got, want, rtol=1e-4, atol=1e-4)
    # mask-sorted row order: same rows, same values (tiles only change which rows they hold)
    order = ops.rulebook_row_order(nbr)
    assert torch.equal(torch.sort(order.long())[0], torch.arange(m, device=cuda))
    got2 = ops.gather_gemm(a, pw, nbr=nbr, act="relu", row_order=order)
    torch.testing.assert_close(got2, want, rtol=1e-4, atol=1e-4)
    # the table permuted into that order and padded to whole tiles (FSFB_NBR_ROW_ORDERED): bit-identical results
    nbr_ro = ops.permute_rulebook(nbr, order)
    assert nbr_ro.shape == (koff, (m + 127) // 128 * 128) and bool((nbr_ro[:, m:] == -1).all())
    got3 = ops.gather_gemm(a, pw, nbr=nbr, act="relu", row_order=order, nbr_ro=nbr_ro)
    assert torch.equal(got3, got2)
    m2 = 70_001                                  # a ragged last tile and a sparse table: padding rows stay without neighbours
    nbr2 = torch.randint(0, m2, (koff, m2), device=cuda, generator=g, dtype=torch.int32)
    nbr2[torch.rand(koff, m2, device=cuda, generator=g) > 0.2] = -1
    a2 = torch.randn(m2, c, device=cuda, generator=g)
    order2 = ops.rulebook_row_order(nbr2)
    ref2 = ops.gather_gemm(a2, pw, nbr=nbr2, act="relu", row_order=order2)
    assert torch.equal(ops.gather_gemm(a2, pw, nbr=nbr2, act="relu", row_order=order2, nbr_ro=ops.permute_rulebook(nbr2, order2)), ref2)


@pytest.mark.parametrize("rows,a_rows,koff,cin,cout,splits,density", [(300, 300, 27, 64, 256, 3, 0.4), (1100, 900, 27, 256, 512, 4, 0.3),
                                                                       (1100, 900, 27, 96, 128, 9, 0.1), (129, 200, 8, 40, 384, 2, 0.6),
                                                                       (700, 700, 27, 512, 512, None, 0.3)])
def test_gather_gemm_splitk(cuda, rows, a_rows, koff, cin, cout, splits, density):
    """Offset-split work units (fsfb_gather_gemm_splitk): partial slabs + reduce/epilogue kernel vs the oracle; includes
    tiles whose split has no active offset at all (sparse masks) and the shape-driven choice of `splits`."""
    rng = np.random.default_rng(rows + koff + cout)
    a = rng.standard_normal((a_rows, cin)).astype(np.float32)
    w = (rng.standard_normal((koff, cout, cin)) / np.sqrt(cin * koff * density)).astype(np.float32)
    nbr = rng.integers(0, a_rows, (koff, rows)).astype(np.int32)
    nbr[rng.random((koff, rows)) > density] = -1
    nbr[: koff // 3, :128] = -1   # first tile: the first split(s) see no offset
    scale = rng.uniform(0.5, 1.5, cout).astype(np.float32)
    shift = rng.standard_normal(cout).astype(np.float32)
    res = rng.standard_normal((rows, cout)).astype(np.float32)
    want = O.gather_gemm(a, w, nbr, norm="affine", norm_w=scale, norm_b=shift, residual=res, act="relu")
    pw = ops.gemm_prepack(T(w, cuda))
    kw = dict(nbr=T(nbr, cuda), norm="affine", norm_w=T(scale, cuda), norm_b=T(shift, cuda), residual=T(res, cuda), act="relu")
    # outputs are O(1..5); a 27 x 512-long 3xTF32 accumulation leaves up to ~3e-5 of the output scale, which shows as a
    # larger relative error only near the ReLU zero crossings: the 1e-4 budget is taken relative to the feature scale
    atol = 4e-5 * float(np.abs(want).max())
    got = ops.gather_gemm(T(a, cuda), pw, splits=splits, **kw).cpu().numpy()
    np.testing.assert_allclose(got, want, rtol=RTOL, atol=atol)
    one = ops.gather_gemm(T(a, cuda), pw, splits=1, **kw).cpu().numpy()
    np.testing.assert_allclose(one, want, rtol=RTOL, atol=atol)


def test_epilogue_vectors_not_cached_by_address(cuda):
    """fsfb_gather_gemm_hv takes host copies of bias / norm vectors; ops caches them per tensor object.  New tensors that
    land on the addresses of freed ones (the caching allocator reuses blocks) must not pick up the old values."""
    rng = np.random.default_rng(5)
    a = rng.standard_normal((500, 64)).astype(np.float32)
    w = (rng.standard_normal((128, 64)) / 8).astype(np.float32)
    pw = ops.gemm_prepack(T(w, cuda))
    ta = T(a, cuda)
    for trial in range(4):
        b = rng.standard_normal(128).astype(np.float32)
        nw = rng.uniform(0.5, 1.5, 128).astype(np.float32)
        nb = rng.standard_normal(128).astype(np.float32)
        tb, tw, th = T(b, cuda), T(nw, cuda), T(nb, cuda)
        got = ops.gather_gemm(ta, pw, bias=tb, norm="affine", norm_w=tw, norm_b=th, act="relu").cpu().numpy()
        np.testing.assert_allclose(got, O.gather_gemm(a, w, bias=b, norm="affine", norm_w=nw, norm_b=nb, act="relu"), rtol=RTOL, atol=ATOL)
        tw.mul_(2.0)   # in-place update of a cached vector
        got = ops.gather_gemm(ta, pw, bias=tb, norm="affine", norm_w=tw, norm_b=th, act="relu").cpu().numpy()
        np.testing.assert_allclose(got, O.gather_gemm(a, w, bias=b, norm="affine", norm_w=2 * nw, norm_b=nb, act="relu"), rtol=RTOL, atol=ATOL)
        del tb, tw, th


@pytest.mark.parametrize("rows,cin,cout,norm,act,res,post", [
    (9000, 128, 128, "ln", "gelu", False, False),      # the SIR point MLP
    (8200, 256, 128, "ln", "gelu", True, False),       # residual before the activation, 8 K chunks
    (10001, 133, 128, "ln", "gelu", False, False),     # unaligned rows: scalar loads, partial last chunk
    (8193, 10, 128, "affine", "relu", False, False),   # one partial chunk (one K step), BatchNorm-as-affine
    (9000, 128, 131, None, None, False, False),        # two column tiles, odd width
    (9000, 128, 33, None, "relu", True, True),         # narrow odd tile, residual after the activation
    (8500, 11, 64, "ln", "relu", False, False),        # 64-column accumulators
    (9100, 32, 146, None, "gelu", True, False),        # two column tiles + residual
    (12345, 181, 128, "ln", None, True, True),
])
def test_linear_row_tile_kernel(cuda, rows, cin, cout, norm, act, res, post):
    """Dense Linear over >= 1024 rows runs csrc/gemm_lin.cu (one CTA per 128-row tile); same oracle, same tolerance."""
    rng = np.random.default_rng(rows + 7 * cin + cout)
    a = rng.standard_normal((rows, cin)).astype(np.float32)
    w = (rng.standard_normal((cout, cin)) / np.sqrt(cin)).astype(np.float32)
    b = rng.standard_normal(cout).astype(np.float32)
    nw = rng.uniform(0.5, 1.5, cout).astype(np.float32)
    nb = rng.standard_normal(cout).astype(np.float32)
    r = rng.standard_normal((rows, cout)).astype(np.float32) if res else None
    want = O.gather_gemm(a, w, bias=b, norm=norm, norm_w=nw, norm_b=nb, eps=1e-3, residual=None if post else r, act=act)
    if post and res:
        want = want + r
    pw = ops.gemm_prepack(T(w, cuda), keep_raw=True)
    kw = dict(bias=T(b, cuda), norm=norm, norm_w=T(nw, cuda) if norm else None, norm_b=T(nb, cuda) if norm else None, eps=1e-3,
              residual=T(r, cuda) if res else None, act=act, residual_post=post)
    got = ops.gather_gemm(T(a, cuda), pw, **kw).cpu().numpy()
    np.testing.assert_allclose(got, want, rtol=RTOL, atol=ATOL)
    # the same rows as a strided view (row stride 192 floats, 16-byte aligned) and into a strided output
    if cin <= 160:
        wide = torch.zeros(rows, 192, device=cuda)
        wide[:, :cin] = T(a, cuda)
        out = torch.full((rows, cout + 5), -7.0, device=cuda)
        ops.gather_gemm(wide[:, :cin], pw, out=out[:, :cout], **kw)
        np.testing.assert_allclose(out[:, :cout].cpu().numpy(), want, rtol=RTOL, atol=ATOL)
        assert float(out[:, cout:].min()) == -7.0 and float(out[:, cout:].max()) == -7.0   # nothing written past the row


@pytest.mark.parametrize("rows,cin,cout,norm,act,res", [
    (3782, 1024, 128, "ln", "gelu", False),     # the refinement heads: 30 row tiles, 32 K chunks -> 8 K ranges
    (3000, 768, 128, "affine", "relu", True),
    (1500, 1000, 33, None, None, False),        # partial last chunk, odd width
    (2048, 1024, 384, "ln", "gelu", True),      # LayerNorm wider than a column tile: fused in the split epilogue
    (1200, 512, 1024, "ln", None, False),       # 1024-wide LayerNorm, two K ranges
    (248, 1024, 128, "ln", "gelu", False),      # the frustum heads: two row tiles
])
def test_linear_k_split(cuda, rows, cin, cout, norm, act, res):
    """Few row tiles x deep K: csrc/gemm_lin.cu runs K ranges as separate CTAs, k_splitk_epilogue sums the slabs (fixed order)."""
    assert ops.linear_k_splits(rows, cin, cout) > 1
    rng = np.random.default_rng(rows + cin + cout)
    a = rng.standard_normal((rows, cin)).astype(np.float32)
    w = (rng.standard_normal((cout, cin)) / np.sqrt(cin)).astype(np.float32)
    b = rng.standard_normal(cout).astype(np.float32)
    nw = rng.uniform(0.5, 1.5, cout).astype(np.float32)
    nb = rng.standard_normal(cout).astype(np.float32)
    r = rng.standard_normal((rows, cout)).astype(np.float32) if res else None
    want = O.gather_gemm(a, w, bias=b, norm=norm, norm_w=nw, norm_b=nb, eps=1e-3, residual=r, act=act)
    pw = ops.gemm_prepack(T(w, cuda), keep_raw=True)
    kw = dict(bias=T(b, cuda), norm=norm, norm_w=T(nw, cuda) if norm else None, norm_b=T(nb, cuda) if norm else None, eps=1e-3,
              residual=T(r, cuda) if res else None, act=act)
    got = ops.gather_gemm(T(a, cuda), pw, **kw)
    np.testing.assert_allclose(got.cpu().numpy(), want, rtol=RTOL, atol=ATOL)
    assert torch.equal(got, ops.gather_gemm(T(a, cuda), pw, **kw))                 # deterministic
    one = ops.gather_gemm(T(a, cuda), pw, splits=1, **kw) if cout <= 256 or norm != "ln" else None
    if one is not None:
        np.testing.assert_allclose(one.cpu().numpy(), want, rtol=RTOL, atol=ATOL)
