"""Dynamic point pooling: oracle invariants (the ones the reference's extractor asserts in-tree)
and GPU parity of fsfb_dynamic_point_pool / the dynamic_point_pool_ext shim / DynamicPointROIExtractor."""
import numpy as np
import pytest

from oracle import fsf_oracle as O


def _scene(n, k, seed, extra=(1.0, 1.0, 1.0)):
    rng = np.random.default_rng(seed)
    pts = np.concatenate([rng.uniform(-50, 50, (n, 2)), rng.uniform(-4, 2, (n, 1))], axis=1).astype(np.float32)
    centers = pts[rng.integers(0, n, k)] + rng.normal(0, 0.3, (k, 3)).astype(np.float32)
    dims = rng.uniform([1.5, 3.0, 1.2], [2.5, 6.0, 2.5], (k, 3))
    rois = np.concatenate([centers, dims, rng.uniform(-np.pi, np.pi, (k, 1))], axis=1).astype(np.float32)
    # dense clusters around some boxes so the per-roi cap matters
    extra_pts = (centers[: max(1, k // 8), None, :] + rng.normal(0, 0.4, (max(1, k // 8), 40, 3))).reshape(-1, 3).astype(np.float32)
    pts = np.concatenate([pts, extra_pts])[rng.permutation(n + extra_pts.shape[0])]
    return pts, rois, list(extra)


def test_oracle_invariants():
    """dynamic_point_roi_extractor.py:84-92 restated on the oracle's output."""
    pts, rois, extra = _scene(5000, 40, 0)
    pi, ri, f = O.dynamic_point_pool(rois, pts, extra, 512, 50000)
    assert pi.size > 0 and (np.diff(ri) >= 0).all()
    roi = rois[ri]
    np.testing.assert_allclose(pts[pi], f[:, :3])
    np.testing.assert_allclose(f[:, 6] + f[:, 9], roi[:, 4], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(f[:, 7] + f[:, 10], roi[:, 3], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(f[:, 8] + f[:, 11], roi[:, 5], rtol=1e-5, atol=1e-5)
    assert (np.abs(f[:, 3]) < roi[:, 4] + extra[0] + 1e-5).all()
    assert (np.abs(f[:, 4]) < roi[:, 3] + extra[1] + 1e-5).all()
    assert (np.abs(f[:, 5]) < roi[:, 5] + extra[2] + 1e-5).all()
    # local coordinates really are the roi frame: rotating back recovers the offset to the centre
    c, s = np.cos(roi[:, 6] + np.pi / 2), np.sin(roi[:, 6] + np.pi / 2)      # local = R(rz + pi/2) shift (mmdet3d 0.x)
    back = np.stack([f[:, 3] * c + f[:, 4] * s, -f[:, 3] * s + f[:, 4] * c], 1)
    np.testing.assert_allclose(back, pts[pi][:, :2] - roi[:, :2], atol=2e-4)
    assert set(np.unique(f[:, 12])) <= {0.0, 1.0} and 0 < f[:, 12].mean() < 1
    # caps
    pi2, ri2, _ = O.dynamic_point_pool(rois, pts, extra, 8, 100)
    assert pi2.size <= 100 and np.bincount(ri2).max() <= 8


@pytest.mark.gpu
@pytest.mark.parametrize("n,k,max_inbox,cap,seed", [(5000, 40, 512, 50000, 1), (60000, 300, 512, 50000, 2), (3000, 9, 16, 50, 3),
                                                    (200, 1, 512, 50000, 4), (20000, 1030, 64, 50000, 5)])
def test_dynamic_point_pool_parity(cuda, n, k, max_inbox, cap, seed):
    import torch
    from fullysparsefusion_b200 import ops
    pts, rois, extra = _scene(n, k, seed)
    want_p, want_r, want_f, amb = O.dynamic_point_pool(rois, pts, extra, max_inbox, cap, margin=1e-4)
    out_p = torch.full((cap,), -1, dtype=torch.int64, device=cuda)
    out_r = torch.full((cap,), -1, dtype=torch.int64, device=cuda)
    out_f = torch.zeros((cap, 13), device=cuda)
    num = ops.dynamic_point_pool(torch.from_numpy(rois).to(cuda), torch.from_numpy(pts).to(cuda), extra, max_inbox, out_p, out_r, out_f)
    m = int(num.item())
    got_p, got_r, got_f = out_p[:m].cpu().numpy(), out_r[:m].cpu().numpy(), out_f[:m].cpu().numpy()
    assert (out_p[m:] == -1).all() and (out_r[m:] == -1).all() and (out_f[m:] == 0).all()   # prefill untouched
    if not amb:   # no point within 1e-4 of a face: ids must match exactly (integer outputs are bit-exact)
        np.testing.assert_array_equal(got_p, want_p)
        np.testing.assert_array_equal(got_r, want_r)
        np.testing.assert_allclose(got_f, want_f, rtol=1e-5, atol=1e-5)
    else:         # cosf/sinf differ from numpy by an ulp: pairs that sit on a face may flip, everything else must agree
        got = set(zip(got_r.tolist(), got_p.tolist()))
        want = set(zip(want_r.tolist(), want_p.tolist()))
        assert (got ^ want) <= amb or max_inbox < 512 or m == cap, sorted(got ^ want)[:5]
    assert (np.diff(got_r) >= 0).all()
    for r in np.unique(got_r)[:20]:
        assert (np.diff(got_p[got_r == r]) > 0).all()


@pytest.mark.gpu
def test_roi_extractor_and_shim(cuda):
    import torch
    from fullysparsefusion_b200 import modules as M
    from fullysparsefusion_b200.shims import dynamic_point_pool_ext
    pts, rois, extra = _scene(8000, 25, 7)
    tp, tr = torch.from_numpy(pts).to(cuda), torch.from_numpy(rois).to(cuda)
    ext = M.DynamicPointROIExtractor(extra_wlh=extra, max_inbox_point=512)
    rois8 = torch.cat([torch.zeros(len(rois), 1, device=cuda), tr], 1)
    inds, roi_inds, info = ext(tp, torch.zeros(len(pts), dtype=torch.int64, device=cuda), rois8)
    want_p, want_r, want_f = O.dynamic_point_pool(rois, pts, extra, 512, 50000)
    assert abs(inds.numel() - want_p.size) <= 2
    assert info["local_xyz"].shape == (inds.numel(), 3) and info["boundary_offset"].shape == (inds.numel(), 6)
    # the shim: same buffers protocol as dynamic_point_pool_op.py:27-32
    op, orr, of = (torch.full((50000,), -1, dtype=torch.int64, device=cuda), torch.full((50000,), -1, dtype=torch.int64, device=cuda),
                   torch.zeros(50000, 13, device=cuda))
    dynamic_point_pool_ext.forward(tr, tp, extra, 512, op, orr, of)
    valid = op >= 0
    assert int(valid.sum()) == inds.numel() and torch.equal(op[valid], inds) and torch.equal(orr[valid], roi_inds)
    # a far-away roi: empty result keeps one fake row (dynamic_point_pool_op.py:36-40)
    far = rois8[:1].clone()
    far[:, 1:4] = 1e4
    i2, r2, _ = ext(tp, torch.zeros(len(pts), dtype=torch.int64, device=cuda), far)
    assert i2.numel() == 1 and int(i2[0]) == -1 and int(r2[0]) == -1


@pytest.mark.gpu
def test_fully_sparse_bbox_head(cuda):
    """Refine-stage head (fsd_bbox_head.py:95-151): point pooling → three SIR blocks over RoI ids → features aligned to
    the RoIs, against the numpy restatement (features 1e-4 relative, the non-empty mask exact)."""
    import torch
    from fullysparsefusion_b200 import modules as M
    from oracle import fsf_oracle_models as OM
    from tests.test_gpu_modules import randomize, sd_np
    pts, rois, extra = _scene(6000, 30, 11)
    rois[-1, :3] = 1e3   # a RoI that pools nothing: stays zero, mask False
    rng = np.random.default_rng(11)
    pts5 = np.concatenate([pts, rng.uniform(0, 1, (len(pts), 2)).astype(np.float32)], 1)
    feats = rng.standard_normal((len(pts), 24)).astype(np.float32)
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda)
    ext = M.DynamicPointROIExtractor(extra_wlh=extra, max_inbox_point=512)
    rois8 = np.concatenate([np.zeros((len(rois), 1), np.float32), rois], 1)
    inds, roi_inds, info = ext(T(pts5), torch.zeros(len(pts), dtype=torch.int64, device=cuda), T(rois8))
    cin = 5 + 24 + 13
    torch.manual_seed(3)
    head = randomize(M.FullySparseBboxHead(num_blocks=3, in_channels=[cin, 5 + 32 + 13, 5 + 32 + 13], feat_channels=[[32, 32]] * 3,
                                           rel_mlp_hidden_dims=[[16, 32]] * 3, rel_mlp_in_channels=[13] * 3, xyz_normalizer=[20, 20, 4],
                                           act="gelu", norm_cfg=dict(type="LN", eps=1e-3), unique_once=True), seed=4).to(cuda)
    ex_pts, ex_feats = T(pts5)[inds], T(feats)[inds]
    got, mask = head(ex_pts, ex_feats, info, roi_inds, T(rois8))
    want, wmask = OM.fully_sparse_bbox_head(ex_pts.cpu().numpy(), ex_feats.cpu().numpy(), info["local_xyz"].cpu().numpy(),
                                            info["boundary_offset"].cpu().numpy(), info["is_in_margin"].cpu().numpy(),
                                            roi_inds.cpu().numpy(), rois8, sd_np(head), 3)
    assert got.shape == (len(rois), 3 * 64)
    np.testing.assert_array_equal(mask.cpu().numpy(), wmask)
    assert not wmask[-1] and wmask[:-1].any()
    np.testing.assert_allclose(got.cpu().numpy(), want, rtol=1e-4, atol=5e-5)


@pytest.mark.gpu
def test_refine_goldens_gpu(cuda):
    """The CUDA path against the goldens recorded from the reference's in-tree refine-stage Python (box decode, extractor
    glue, RoI alignment)."""
    import torch
    from fullysparsefusion_b200 import modules as M, ops
    from tests.conftest import load_golden
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda)
    g = load_golden("box_decode")
    np.testing.assert_allclose(ops.decode_boxes(T(g["reg"]), T(g["base"])).cpu().numpy(), g["rois"], rtol=2e-6, atol=2e-6)
    g = load_golden("roi_extractor")
    ext = M.DynamicPointROIExtractor(extra_wlh=[1.0, 1.0, 1.0], max_inbox_point=512)
    inds, roi_inds, info = ext(T(g["points"]), None, T(g["rois"]))
    assert np.array_equal(inds.cpu().numpy(), g["inds"]) and np.array_equal(roi_inds.cpu().numpy(), g["roi_inds"])
    np.testing.assert_allclose(info["local_xyz"].cpu().numpy(), g["local_xyz"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(info["boundary_offset"].cpu().numpy(), g["boundary_offset"], rtol=1e-5, atol=1e-5)
    np.testing.assert_array_equal(info["is_in_margin"].cpu().numpy(), g["is_in_margin"])
    g = load_golden("roi_align")
    new, mask = M.FullySparseBboxHead.align(T(g["feats"]), T(g["out_coors"]), int(g["num_rois"]))
    assert np.array_equal(mask.cpu().numpy(), g["mask"])
    np.testing.assert_array_equal(new.cpu().numpy(), g["aligned"])
