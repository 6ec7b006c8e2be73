"""GPU parity of the registry-level modules (DynamicScatterVFE, SIRLayer/SIR, SimpleSparseUNet, neck,
VoteSegHead) against the numpy model oracle with the same weights.  1e-4 relative on features; voxel
coordinates / inverse indices bit-exact."""
import numpy as np
import pytest
import torch
from torch import nn

from fullysparsefusion_b200 import modules as M
from fullysparsefusion_b200 import ops, synth
from oracle import fsf_oracle as O
from oracle import fsf_oracle_models as OM
from tests.conftest import load_golden

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-4, 3e-5
BN = dict(type="naiveSyncBN1d", eps=1e-3, momentum=0.01)
LN = dict(type="LN", eps=1e-3)


def T(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def randomize(mod: nn.Module, seed=0):
    g = torch.Generator().manual_seed(seed)
    for m in mod.modules():
        if isinstance(m, nn.BatchNorm1d):
            m.running_mean.copy_(torch.randn(m.num_features, generator=g) * 0.3)
            m.running_var.copy_(torch.rand(m.num_features, generator=g) * 1.5 + 0.5)
        if isinstance(m, (nn.BatchNorm1d, nn.LayerNorm)):
            m.weight.data.copy_(torch.rand(m.weight.shape, generator=g) + 0.5)
            m.bias.data.copy_(torch.randn(m.bias.shape, generator=g) * 0.2)
    return mod.eval()


def sd_np(mod):
    return {k: v.detach().cpu().numpy() for k, v in mod.state_dict().items()}


def _points_coors(n, seed, sweeps=1):
    pts = synth.ring_points(n, sweeps=sweeps, seed=seed)[:, :5]
    c = O.voxelize(pts, synth.NUSC_VOXEL, synth.NUSC_RANGE, floor_mode=0).astype(np.int64)
    return pts, np.concatenate([np.zeros((n, 1), np.int64), c], 1)


def test_dynamic_scatter_vfe(cuda):
    pts, coors = _points_coors(20000, 3)
    vfe = randomize(M.DynamicScatterVFE(in_channels=5, feat_channels=[64, 64], with_cluster_center=True, with_voxel_center=True,
                                        voxel_size=synth.NUSC_VOXEL, point_cloud_range=synth.NUSC_RANGE, norm_cfg=BN,
                                        unique_once=True)).to(cuda)
    vf, vc, inv = vfe(T(pts, cuda), T(coors, cuda), return_inv=True)
    w_vf, w_vc, w_inv = OM.dynamic_scatter_vfe(pts, coors, sd_np(vfe), synth.NUSC_VOXEL, synth.NUSC_RANGE)
    assert np.array_equal(vc.cpu().numpy(), w_vc) and np.array_equal(inv.cpu().numpy(), w_inv)
    np.testing.assert_allclose(vf.cpu().numpy(), w_vf, rtol=RTOL, atol=ATOL)


def test_sir_backbone(cuda):
    rng = np.random.default_rng(0)
    n, k = 6000, 90
    ids = rng.integers(0, k, n)
    coors = np.stack([rng.integers(0, 3, n) * 0 + ids % 3, np.zeros(n, np.int64), ids], 1).astype(np.int64)
    pts = synth.ring_points(n, seed=5)[:, :5]
    feats = rng.standard_normal((n, 43)).astype(np.float32)
    f_cluster = rng.standard_normal((n, 3)).astype(np.float32)
    sir = randomize(M.SIR(num_blocks=3, in_channels=[48, 37, 37], feat_channels=[[32, 32]] * 3, rel_mlp_hidden_dims=[[16, 32]] * 3,
                          norm_cfg=LN, mode="max", xyz_normalizer=[20, 20, 4], act="gelu", unique_once=True)).to(cuda)
    out, cl, oc = sir(T(pts, cuda), T(feats, cuda), T(coors, cuda), T(f_cluster, cuda))
    w_out, w_cl, w_oc = OM.sir(pts, feats, coors, f_cluster, sd_np(sir), 3, [20, 20, 4])
    assert np.array_equal(oc.cpu().numpy(), w_oc)
    assert cl.shape == (len(w_oc), 192)
    np.testing.assert_allclose(cl.cpu().numpy(), w_cl, rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(out.cpu().numpy(), w_out, rtol=RTOL, atol=ATOL)


def test_simple_sparse_unet(cuda):
    """5-stage U-Net (the stock topology with narrower channels so the float64 oracle stays fast)."""
    pts, coors = _points_coors(4000, 9)
    uniq = O.unique_rows(coors)[0]
    rng = np.random.default_rng(1)
    feats = rng.standard_normal((len(uniq), 16)).astype(np.float32)
    enc = ((32,), (32, 32, 32), (32, 32, 32), (48, 48, 48), (64, 64, 64))
    encp = ((1,), (1, 1, 1), (1, 1, 1), ((0, 1, 1), 1, 1), (1, 1, 1))
    dec = ((64, 64, 48), (48, 48, 32), (32, 32, 32), (32, 32, 32), (32, 32, 32))
    net = randomize(M.SimpleSparseUNet(in_channels=16, sparse_shape=[40, 512, 512], norm_cfg=BN, base_channels=16,
                                       output_channels=32, encoder_channels=enc, encoder_paddings=encp, decoder_channels=dec,
                                       decoder_paddings=((1, 1), (1, 0), (1, 0), (0, 0), (0, 1)))).to(cuda)
    n_convs = sum(isinstance(m, M.SparseConvModule) for m in net.modules())
    assert n_convs == 34                                  # ~34 sparse convolutions
    out = net(dict(voxel_feats=T(feats, cuda), voxel_coors=T(uniq, cuda), batch_size=1))[0]["voxel_feats"]
    want, rb, levels = OM.simple_sparse_unet(feats, uniq, sd_np(net), [40, 512, 512], enc, encp, dec)
    assert out.shape == want.shape == (len(uniq), 32)
    got_rb, got_levels = net.build_rulebooks(T(uniq.astype(np.int32), cuda),
                                             ops.unique_rows(T(uniq.astype(np.int32), cuda), lo=[0] * 4, ext=[1, 40, 512, 512],
                                                             return_index=True)[3], 1)
    for key in rb:                                         # rulebooks bit-exact at every level
        assert np.array_equal(got_rb[key].nbr.cpu().numpy(), rb[key]), key
        if got_rb[key].order is not None:                  # a permutation, sorted by offset mask
            order = got_rb[key].order.cpu().numpy()
            assert np.array_equal(np.sort(order), np.arange(rb[key].shape[1]))
            mask = ((rb[key] >= 0).astype(np.int64) << np.arange(27)[:, None]).sum(0)
            assert np.array_equal(order, np.argsort(mask, kind="stable"))
    for a, b in zip(got_levels, levels):
        assert np.array_equal(a["coors"].cpu().numpy(), b["coors"])
    np.testing.assert_allclose(out.cpu().numpy(), want, rtol=2e-4, atol=1e-4)   # 34 chained layers


def test_neck_golden(cuda):
    g = load_golden("neck")
    neck = M.Voxel2PointScatterNeck(point_cloud_range=[-51.2, -51.2, -5.0, 51.2, 51.2, 3.0], voxel_size=[0.2, 0.2, 0.2])
    out, mask = neck(T(g["points"], cuda), T(g["coors"], cuda), T(g["voxel_feats"], cuda), T(g["voxel2point_inds"], cuda), -1)
    assert np.array_equal(mask.cpu().numpy(), g["mask"])
    assert np.array_equal(out.cpu().numpy(), g["out"])      # the reference module's own output, bit-exact


def test_vote_seg_head(cuda):
    rng = np.random.default_rng(2)
    x = rng.standard_normal((5000, 131)).astype(np.float32)
    head = randomize(M.VoteSegHead(in_channel=131, num_classes=10, hidden_dims=[128, 128])).to(cuda)
    logits, votes = head(T(x, cuda))
    w_logits, w_votes = OM.vote_seg_head(x, sd_np(head))
    assert logits.shape == (5000, 11) and votes.shape == (5000, 33)
    np.testing.assert_allclose(logits.cpu().numpy(), w_logits, rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(votes.cpu().numpy(), w_votes, rtol=RTOL, atol=ATOL)
    g = load_golden("vote_decode")
    assert np.array_equal(head.decode_vote_targets(T(g["preds"], cuda)).cpu().numpy(), g["offsets"])


def test_compact_indices(cuda):
    rng = np.random.default_rng(4)
    for n in (0, 1, 2047, 2049, 300000):
        mask = rng.random(n) < 0.3
        got = ops.compact_indices(T(mask, cuda))
        assert np.array_equal(got.cpu().numpy(), np.flatnonzero(mask).astype(np.int32))


def test_cluster_head_reference_golden(cuda):
    """fsf.SparseClusterHeadV2 with the state dict of the reference's own head: same logits / regressions (1e-4)."""
    from tests.test_oracle_golden import _v2_head_from_golden
    g, head, sd = _v2_head_from_golden()
    head.load_state_dict(sd, strict=True)
    head = head.eval().to(cuda)
    with torch.no_grad():
        out = head(torch.from_numpy(g["x"]).to(cuda))
    np.testing.assert_allclose(out["cls_logits"][0].cpu().numpy(), g["cls"], rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(out["reg_preds"][0].cpu().numpy(), g["reg"], rtol=1e-4, atol=2e-5)


def test_vote_seg_head_reference_golden(cuda):
    """modules.VoteSegHead with the state dict of the reference's own VoteSegHead (BN running statistics folded into the fused
    Linear epilogue): same logits and votes (1e-4)."""
    from tests.test_oracle_golden import _seg_head_from_golden
    g, head, sd = _seg_head_from_golden()
    head.load_state_dict(sd, strict=True)
    head = head.eval().to(cuda)
    with torch.no_grad():
        logits, votes = head(torch.from_numpy(g["x"]).to(cuda))
    np.testing.assert_allclose(logits.cpu().numpy(), g["logits"], rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(votes.cpu().numpy(), g["votes"], rtol=1e-4, atol=2e-5)
