"""The reference's OWN SIR.forward (models/backbones/sir.py:65-85) and FullySparseBboxHead.forward (models/roi_heads/bbox_heads/
fsd_bbox_head.py:95-151) were run by `tools/make_golden.py wiring` with the registry types 'SIRLayer' / 'DynamicClusterVFE' built
from the repo's block classes (CPU stand-in forward).  Here the repo's modules.SIR / modules.FullySparseBboxHead run the same
inputs and parameters on the GPU: the in-tree control flow around the un-vendored block (row concatenation, feature order, the
-1 group, RoI alignment) is pinned to the reference's."""
import json
import os

import numpy as np
import pytest
import torch

from tests.conftest import GOLDEN, load_golden

pytestmark = pytest.mark.gpu


def _sd(g, prefix):
    return {k[len(prefix):].replace("__", "."): torch.from_numpy(v) for k, v in g.items() if k.startswith(prefix)}


def test_sir_matches_the_reference_forward(cuda):
    from fullysparsefusion_b200 import modules as M
    g = load_golden("wiring_sir")
    kw = json.load(open(os.path.join(GOLDEN, "wiring_kwargs.json")))["sir"]
    net = M.SIR(**kw)
    net.load_state_dict(_sd(g, "sir_sd__"), strict=True)      # the reference module's own state dict (blocks are the repo's)
    net = net.to(cuda).eval()
    T = lambda k: torch.from_numpy(g[k]).to(cuda)
    with torch.no_grad():
        out_feats, cluster_feats, out_coors = net(T("points"), T("features"), T("coors"), T("f_cluster"))
    assert np.array_equal(out_coors.cpu().numpy().astype(np.int64), g["out_coors"])
    np.testing.assert_allclose(cluster_feats.cpu().numpy(), g["cluster_feats"], rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(out_feats.cpu().numpy(), g["out_feats"], rtol=1e-4, atol=2e-5)


def test_roi_head_matches_the_reference_forward(cuda):
    from fullysparsefusion_b200 import modules as M
    g = load_golden("wiring_roi_head")
    kw = json.load(open(os.path.join(GOLDEN, "wiring_kwargs.json")))["head"]
    head = M.FullySparseBboxHead(**kw)
    head.load_state_dict(_sd(g, "head_sd__"), strict=True)
    head = head.to(cuda).eval()
    T = lambda k: torch.from_numpy(g[k]).to(cuda)
    info = dict(local_xyz=T("local_xyz"), boundary_offset=T("boundary_offset"), is_in_margin=T("is_in_margin"))
    with torch.no_grad():
        feats, nonempty = head(T("pts_xyz"), T("pts_features"), info, T("roi_inds"), T("rois"))
    assert np.array_equal(nonempty.cpu().numpy(), g["nonempty"])
    np.testing.assert_allclose(feats.cpu().numpy(), g["roi_feats"], rtol=1e-4, atol=2e-5)


def test_dynamic_cluster_vfe_call_convention(cuda):
    """fsd_bbox_head.py:135,140: block(in_feats, roi_inds [P] i64, f_cluster, unq_inv_once=, new_coors_once=) — 1-d ids in, 1-d i64
    group ids out for the last block, (point feats, group feats) for the others."""
    from fullysparsefusion_b200 import modules as M
    torch.manual_seed(0)
    blk = M.DynamicClusterVFE(in_channels=20, feat_channels=[32, 32], with_rel_mlp=True, rel_mlp_hidden_dims=[16, 32], rel_mlp_in_channel=13,
                              norm_cfg=dict(type="LN", eps=1e-3, momentum=0.01), mode="max", return_point_feats=False, rel_dist_scaler=10.0,
                              fusion="cat", pos_fusion="mul", xyz_normalizer=[20, 20, 4], cat_voxel_feats=True, act="gelu", dropout=0,
                              with_distance=False, with_cluster_center=False, with_voxel_center=False, voxel_size=[0.1, 0.1, 0.1],
                              point_cloud_range=[-74.88, -74.88, -2, 74.88, 74.88, 4], fusion_layer=None, return_inv=False).to(cuda).eval()
    x = torch.randn(300, 20, device=cuda)
    ids = torch.randint(-1, 9, (300,), device=cuda)
    with torch.no_grad():
        feats, coors = blk(x, ids, torch.randn(300, 13, device=cuda), unq_inv_once=None, new_coors_once=None)
    assert coors.dim() == 1 and coors.dtype == torch.int64 and torch.equal(coors, torch.unique(ids))
    assert feats.shape == (coors.numel(), 64)
