"""GPU parity: the CUDA path (through the C-ABI in libfsf_b200.so) against the CPU oracle and the
reference-generated goldens, on the same seeded inputs.  Bit-exact for coordinates, ranks, CSR,
argmax and in-group indices; 1e-4 relative for fp32 sums/means (north_star tolerance)."""
import numpy as np
import pytest
import torch

from fullysparsefusion_b200 import ops, synth
from oracle import fsf_oracle as O
from tests.conftest import load_golden
from tests.test_oracle_golden import _proj_inputs, _scatter_feat

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-4, 1e-5


def T(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


# ---- a1 voxelize -----------------------------------------------------------------------------
@pytest.mark.parametrize("tag", ["voxel_nusc", "voxel_pre"])
@pytest.mark.parametrize("mode,key", [(0, "coors_floor"), (1, "coors_divfloor")])
def test_voxelize_golden(cuda, tag, mode, key):
    g = load_golden(tag)
    rng, vs = g["pc_range"].tolist(), g["voxel_size"].tolist()
    got = ops.voxelize(T(g["points"], cuda), vs, rng, floor_mode=mode, grid=[4096] * 3).cpu().numpy()
    ok = np.all(g[key] >= 0, axis=1)
    assert np.array_equal(got[ok], g[key][ok].astype(np.int32))
    assert np.all(got[~ok] == -1)


@pytest.mark.parametrize("n,gen", [(0, "ring"), (1, "ring"), (34000, "ring"), (300000, "ring"), (100003, "uniform")])
@pytest.mark.parametrize("mode", [0, 1])
def test_voxelize_oracle(cuda, n, gen, mode):
    pts = synth.ring_points(n, sweeps=10 if n > 100000 else 1, seed=n) if gen == "ring" and n else synth.uniform_points(n, seed=3)
    if n:  # out-of-range and non-finite rows
        pts[::97, 0] = 60.0
        pts[5 % n, 1] = np.nan
        pts[7 % n, 2] = -np.inf
    got = ops.voxelize(T(pts, cuda), synth.NUSC_VOXEL, synth.NUSC_RANGE, floor_mode=mode).cpu().numpy()
    want = O.voxelize(pts, synth.NUSC_VOXEL, synth.NUSC_RANGE, floor_mode=mode)
    assert np.array_equal(got, want)


# ---- a2 ranking == torch.unique(dim=0) ---------------------------------------------------------
def _coors(n, seed, sweeps=1, vs=synth.NUSC_VOXEL, batches=1):
    pts = synth.ring_points(n, sweeps=sweeps, seed=seed)
    c = O.voxelize(pts, vs, synth.NUSC_RANGE, floor_mode=1).astype(np.int64)
    b = np.random.default_rng(seed).integers(0, batches, (n, 1))
    return np.concatenate([b, c], 1)


@pytest.mark.parametrize("tag", ["scatter_c1", "scatter_vox", "scatter_odd", "scatter_ties"])
def test_unique_rows_golden(cuda, tag):
    g = load_golden(tag)
    uniq, inv, cnt = ops.unique_rows(T(g["coors"].astype(np.int64), cuda), return_counts=True)
    assert np.array_equal(uniq.cpu().numpy(), g["new_coors"])
    assert np.array_equal(inv.cpu().numpy(), g["unq_inv"])
    assert np.array_equal(cnt.cpu().numpy(), np.bincount(g["unq_inv"]))


@pytest.mark.parametrize("n,sweeps,batches,dtype", [(1, 1, 1, np.int64), (777, 1, 2, np.int32), (34000, 1, 1, np.int64),
                                                    (300000, 10, 2, np.int64)])
def test_unique_rows_oracle(cuda, n, sweeps, batches, dtype):
    rows = _coors(n, 21, sweeps, batches=batches).astype(dtype)
    u, inv, cnt = O.unique_rows(rows)
    # bounds known (voxel grid) and bounds discovered (min/max pass) must agree
    for kw in (dict(), dict(lo=[0, 0, 0, 0], ext=[batches, 40, 512, 512])):
        gu, ginv, gcnt = ops.unique_rows(T(rows, cuda), return_counts=True, **kw)
        assert gu.dtype == torch.from_numpy(rows).dtype
        assert np.array_equal(gu.cpu().numpy(), u)
        assert np.array_equal(ginv.cpu().numpy(), inv)
        assert np.array_equal(gcnt.cpu().numpy(), cnt)


def test_unique_rows_empty_and_negative(cuda):
    u, inv, cnt = ops.unique_rows(torch.zeros((0, 4), dtype=torch.int64, device=cuda), return_counts=True)
    assert u.shape == (0, 4) and inv.numel() == 0 and cnt.numel() == 0
    rows = np.array([[0, -3, 5], [0, -3, 5], [-1, 7, 0], [2, 0, 0], [0, -3, 4]], np.int64)
    u, inv, _ = O.unique_rows(rows)
    gu, ginv, _ = ops.unique_rows(T(rows, cuda))
    assert np.array_equal(gu.cpu().numpy(), u) and np.array_equal(ginv.cpu().numpy(), inv)


# ---- CSR + segmented reductions == torch_scatter -----------------------------------------------
def _check_csr(csr, index, m):
    off, perm, seg = csr.offsets.cpu().numpy(), csr.perm.cpu().numpy(), csr.seg.cpu().numpy()
    order = np.argsort(np.where((index < 0) | (index >= m), m, index), kind="stable")
    assert np.array_equal(perm, order.astype(np.int32))          # stable: ascending rows inside a segment
    cnt = np.bincount(index[(index >= 0) & (index < m)], minlength=m)
    assert np.array_equal(off, np.r_[0, np.cumsum(cnt)].astype(np.int32))
    assert np.array_equal(seg[: off[-1]], index[order][: off[-1]].astype(np.int32))


@pytest.mark.parametrize("tag", ["scatter_c1", "scatter_vox", "scatter_odd", "scatter_ties"])
def test_scatter_golden(cuda, tag):
    g = load_golden(tag)
    feat = _scatter_feat(g)
    inv = g["unq_inv"].astype(np.int64)
    m = int(inv.max()) + 1
    csr = ops.build_csr(T(inv, cuda), m)
    _check_csr(csr, inv, m)
    f = T(feat, cuda)
    out, arg = ops.segment_reduce(f, csr, "max", return_argmax=True)
    assert np.array_equal(out.cpu().numpy(), g["out_max"])
    assert np.array_equal(arg.cpu().numpy(), g["argmax"])
    assert np.array_equal(ops.segment_reduce(f, csr, "max").cpu().numpy(), g["out_max"])
    np.testing.assert_allclose(ops.segment_reduce(f, csr, "mean").cpu().numpy(), g["out_avg"], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(ops.segment_reduce(f, csr, "sum").cpu().numpy(), g["out_sum"], rtol=RTOL, atol=1e-4)


@pytest.mark.parametrize("n,c,m,kind", [
    (1, 1, 1, "ids"), (33, 3, 5, "ids"), (5000, 5, 40, "ids"), (10000, 128, 32, "ids"), (4097, 131, 1000, "ids"),
    (20000, 64, 3, "ids"), (34000, 11, None, "vox"), (34000, 33, None, "vox"), (60000, 16, None, "vox"),
    (30000, 180, 250, "ids"), (3000, 768, 17, "ids"), (100000, 4, 50000, "holes"),
])
def test_scatter_oracle(cuda, n, c, m, kind):
    rng = np.random.default_rng(n + c)
    feat = rng.standard_normal((n, c)).astype(np.float32)
    feat[rng.random((n, c)) < 0.2] = np.float32(0.5)  # ties
    if kind == "vox":
        _, index, _ = O.unique_rows(_coors(n, 5))
        m = int(index.max()) + 1
    elif kind == "holes":  # empty segments and dropped rows (index < 0)
        index = rng.integers(0, m, n) * 2 % m
        index[rng.random(n) < 0.05] = -1
    else:
        index = rng.integers(0, m, n)
    index = index.astype(np.int64)
    csr = ops.build_csr(T(index, cuda), m)
    _check_csr(csr, index, m)
    f = T(feat, cuda)
    w_max, w_arg = O.scatter_max(feat, index, m)
    g_max, g_arg = ops.segment_reduce(f, csr, "max", return_argmax=True)
    assert np.array_equal(g_max.cpu().numpy(), w_max)
    if kind == "holes":  # argmax of rows the oracle numbers in the full array
        assert np.array_equal(g_arg.cpu().numpy(), w_arg)
    else:
        assert np.array_equal(g_arg.cpu().numpy(), w_arg)
    np.testing.assert_allclose(ops.segment_reduce(f, csr, "mean").cpu().numpy(), O.scatter_mean(feat, index, m), rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(ops.segment_reduce(f, csr, "sum").cpu().numpy(), O.scatter_sum(feat, index, m), rtol=RTOL, atol=1e-4)


def test_scatter_strided_and_int32_index(cuda):
    rng = np.random.default_rng(0)
    wide = rng.standard_normal((9000, 40)).astype(np.float32)
    index = rng.integers(0, 77, 9000).astype(np.int32)
    csr = ops.build_csr(T(index, cuda), 77)
    view = T(wide, cuda)[:, 4:24]  # row stride 40, 20 channels
    got = ops.segment_reduce(view, csr, "max")
    assert np.array_equal(got.cpu().numpy(), O.scatter_max(wide[:, 4:24], index, 77)[0])


def test_scatter_large_properties(cuda):
    """Full-size (C3/C5) checks through size-independent properties: sum linearity and the
    checksum of segment sums == checksum of the input; max idempotence via gather."""
    n, c = 1_000_000, 64
    g = torch.Generator(device=cuda).manual_seed(0)
    feat = torch.randn(n, c, device=cuda, generator=g)
    index = torch.randint(0, 200_000, (n,), device=cuda, generator=g)
    csr = ops.build_csr(index, 200_000)
    s = ops.segment_reduce(feat, csr, "sum")
    torch.testing.assert_close(s.double().sum(0), feat.double().sum(0), rtol=1e-6, atol=1e-3)
    s2 = ops.segment_reduce(feat * 2 + 1, csr, "sum")
    cnt = torch.bincount(index, minlength=200_000).float()[:, None]
    torch.testing.assert_close(s2, 2 * s + cnt, rtol=1e-4, atol=1e-3)
    mx, arg = ops.segment_reduce(feat, csr, "max", return_argmax=True)
    live = cnt[:, 0] > 0
    assert torch.equal(feat.gather(0, arg[live].clamp(max=n - 1)), mx[live])   # argmax points at the max
    assert torch.equal(index[arg[live][:, 0]], torch.nonzero(live)[:, 0])      # ... inside the right segment
    back = ops.gather_rows(mx, index)
    assert bool((back >= feat).all())
    mx2 = ops.segment_reduce(back, csr, "max")
    assert torch.equal(mx2, mx)                                                # idempotent


# ---- gather / in-group --------------------------------------------------------------------------
@pytest.mark.parametrize("c", [3, 64, 128, 131])
def test_gather_rows(cuda, c):
    rng = np.random.default_rng(c)
    src = rng.standard_normal((1000, c)).astype(np.float32)
    idx = rng.integers(-1, 1000, 20000).astype(np.int64)
    got = ops.gather_rows(T(src, cuda), T(idx, cuda), fill=-7.0)
    assert np.array_equal(got.cpu().numpy(), O.gather_rows(src, idx, fill=-7.0))
    got32 = ops.gather_rows(T(src, cuda), T(idx.astype(np.int32), cuda), fill=-7.0)
    assert torch.equal(got, got32)


def test_ingroup_golden(cuda):
    g = load_golden("ingroup")
    got = ops.ingroup_indices(T(g["group"].astype(np.int64), cuda))
    assert np.array_equal(got.cpu().numpy(), g["inner"])


@pytest.mark.parametrize("n,m", [(1, 1), (5000, 3), (100000, 2049), (300000, 150000)])
def test_ingroup_oracle(cuda, n, m):
    grp = np.random.default_rng(n).integers(0, m, n).astype(np.int64)
    got = ops.ingroup_indices(T(grp, cuda), m)
    assert np.array_equal(got.cpu().numpy(), O.ingroup_indices(grp))


# ---- a7 + a8 projection + nearest sampling ------------------------------------------------------
@pytest.mark.parametrize("tag", ["projection_small", "projection_nusc"])
def test_projection_golden(cuda, tag):
    g = load_golden(tag)
    pts, l2i, mask, H, W = _proj_inputs(g)
    ids = ops.project_sample(T(pts, cuda), T(l2i, cuda), T(mask, cuda)).cpu().numpy()
    ref_ids = g["ids"].astype(np.int64)
    flips = np.any(ids != ref_ids, axis=(1, 2)).mean()
    assert flips <= 2e-3, f"flip rate vs reference grid_sample {flips}"   # texel-boundary rounding only
    want = O.points_in_mask(pts, mask, l2i)
    assert np.array_equal(ids, want)                                       # bit-exact vs the oracle


@pytest.mark.parametrize("n,sweeps,dtype", [(0, 1, np.uint8), (31, 1, np.uint8), (34000, 1, np.uint8), (300000, 10, np.uint8),
                                            (20000, 1, np.int32)])
def test_projection_oracle(cuda, n, sweeps, dtype):
    H, W = (900, 1600) if n >= 34000 else (90, 160)
    pts = synth.ring_points(max(n, 1), sweeps=sweeps, seed=9)[:n, :3]
    l2i = synth.lidar2img(6, H, W)
    mask = synth.mask_planes(6, 10, H, W, seed=4, overlap=True, dtype=dtype)
    want = O.points_in_mask(pts, mask, l2i)
    dpts, dl2i, dmask = T(pts.reshape(-1, 3), cuda), T(l2i, cuda), T(mask, cuda)
    ids = ops.project_sample(dpts, dl2i, dmask)
    assert ids.shape == (n, 6, 10) and ids.dtype == torch.int64
    assert np.array_equal(ids.cpu().numpy(), want)
    ids_sel, cam, fg, ov = ops.project_sample_select(dpts, dl2i, dmask, want_overlap=True)
    w_ids, w_cam, w_fg, w_ov = O.cam_select(want)
    assert np.array_equal(ids_sel.cpu().numpy(), w_ids)
    assert np.array_equal(cam.cpu().numpy(), w_cam)
    assert np.array_equal(fg.cpu().numpy().astype(bool), w_fg)
    assert np.array_equal(ov.cpu().numpy(), w_ov)
    if n:
        assert w_fg.mean() > 0.01  # the synthetic scene does hit masks


# ---- a14 connected components -------------------------------------------------------------------
@pytest.mark.parametrize("tag", ["ccl_small", "ccl_mid", "ccl_batch"])
def test_ccl_golden(cuda, tag):
    g = load_golden(tag)
    pts, b = T(g["points"], cuda), T(g["batch_idx"], cuda)
    lab, cnt = ops.connected_components(pts, b if tag == "ccl_batch" else None, float(g["dist"]), return_count=True)
    assert np.array_equal(lab.cpu().numpy(), g["labels"])          # bit-exact vs the reference's scipy path
    assert int(cnt) == int(g["labels"].max()) + 1


@pytest.mark.parametrize("m,batches,dist", [(0, 1, 0.5), (1, 1, 0.5), (257, 1, 0.3), (10000, 1, 0.6), (30000, 4, 0.4),
                                            (5000, 3, 2.0)])
def test_ccl_oracle(cuda, m, batches, dist):
    pts, b = synth.cluster_points(max(m, 1), seed=m, batches=batches)
    pts, b = pts[:m], b[:m]
    single = ops.connected_components(T(pts.reshape(-1, 3), cuda), None, dist).cpu().numpy()
    assert np.array_equal(single, O.connected_components_single_batch(pts, dist))
    if m:
        # batch ids in arbitrary (unsorted) order and sorted order
        for bb in (b, np.sort(b)):
            got = ops.connected_components(T(pts, cuda), T(bb, cuda), dist).cpu().numpy()
            assert np.array_equal(got, O.connected_components(pts, bb, dist))
        assert len(np.unique(single)) == single.max() + 1              # the reference's own assert (:42)


@pytest.mark.parametrize("n,m,kind", [(30000, None, "vox"), (20000, 3500, "ids"), (9000, 6000, "holes"), (5, 3, "ids")])
def test_scatter_padded_131_rows(cuda, n, m, kind):
    """The 131-channel point features live in 16-byte padded rows (ops.empty_rows → 132 floats): short segments take the
    dedicated k_segreduce_small_132 kernel (33rd float4 of four rows in one load).  Same oracle, same exactness."""
    rng = np.random.default_rng(n)
    feat = rng.standard_normal((n, 131)).astype(np.float32)
    feat[rng.random((n, 131)) < 0.2] = np.float32(0.5)
    if kind == "vox":
        _, index, _ = O.unique_rows(_coors(n, 5))
        m = int(index.max()) + 1
    elif kind == "holes":
        index = rng.integers(0, m, n) * 2 % m
        index[rng.random(n) < 0.05] = -1
    else:
        index = rng.integers(0, m, n)
    index = index.astype(np.int64)
    csr = ops.build_csr(T(index, cuda), m)
    f = ops.empty_rows(n, 131, cuda)
    f.copy_(T(feat, cuda))
    assert f.stride(0) == 132
    assert np.array_equal(ops.segment_reduce(f, csr, "max").cpu().numpy(), O.scatter_max(feat, index, m)[0])
    np.testing.assert_allclose(ops.segment_reduce(f, csr, "mean").cpu().numpy(), O.scatter_mean(feat, index, m), rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(ops.segment_reduce(f, csr, "sum").cpu().numpy(), O.scatter_sum(feat, index, m), rtol=RTOL, atol=1e-4)
    g_max, g_arg = ops.segment_reduce(f, csr, "max", return_argmax=True)   # argmax keeps the generic kernel
    w_max, w_arg = O.scatter_max(feat, index, m)
    assert np.array_equal(g_max.cpu().numpy(), w_max) and np.array_equal(g_arg.cpu().numpy(), w_arg)
