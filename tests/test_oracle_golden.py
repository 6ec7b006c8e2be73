"""Pin the CPU oracle against outputs of the REFERENCE's own Python (tests/golden/*.npz, written by
tools/make_golden.py in the build container).  CPU only; no CUDA, no /root/reference at run time."""
import hashlib

import numpy as np
import pytest

from fullysparsefusion_b200 import synth
from oracle import fsf_oracle as O
from tests.conftest import load_golden


def _scatter_feat(g):
    if g["feat"].size:
        return g["feat"]
    import torch

    n, c = (int(v) for v in g["feat_shape"])
    feat = torch.randn(n, c, generator=torch.Generator().manual_seed(int(g["feat_seed"]))).numpy()
    sha = np.frombuffer(hashlib.sha256(feat.tobytes()).digest(), dtype=np.uint8)
    assert np.array_equal(sha, g["feat_sha"]), "torch.randn stream changed: regenerate goldens"
    return feat


@pytest.mark.parametrize("tag", ["scatter_c1", "scatter_vox", "scatter_odd", "scatter_ties"])
def test_scatter_v2(tag):
    g = load_golden(tag)
    feat = _scatter_feat(g)
    coors = g["coors"].astype(np.int64)
    uniq, inv, _ = O.unique_rows(coors)
    assert np.array_equal(uniq, g["new_coors"])          # torch.unique(dim=0) order, bit-exact
    assert np.array_equal(inv, g["unq_inv"])
    out, arg = O.scatter_max(feat, inv)
    assert np.array_equal(out, g["out_max"])             # max is exact
    assert np.array_equal(arg, g["argmax"])              # first-max-wins
    np.testing.assert_allclose(O.scatter_mean(feat, inv), g["out_avg"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(O.scatter_sum(feat, inv), g["out_sum"], rtol=1e-4, atol=1e-4)
    f2, c2, i2 = O.scatter_v2(feat, coors, "max")
    assert np.array_equal(f2, g["out_max"]) and np.array_equal(c2, g["new_coors"]) and np.array_equal(i2, inv)


@pytest.mark.parametrize("tag", ["voxel_nusc", "voxel_pre"])
def test_voxel_coords(tag):
    g = load_golden(tag)
    rng, vs = g["pc_range"].tolist(), g["voxel_size"].tolist()
    big = [4096, 4096, 4096]  # goldens hold raw coordinates (no range rejection)
    c1 = O.voxelize(g["points"], vs, rng, floor_mode=1, grid=big)
    c0 = O.voxelize(g["points"], vs, rng, floor_mode=0, grid=big)
    ok1 = np.all(g["coors_divfloor"] >= 0, axis=1)
    ok0 = np.all(g["coors_floor"] >= 0, axis=1)
    assert np.array_equal(c1[ok1], g["coors_divfloor"][ok1])
    assert np.array_equal(c0[ok0], g["coors_floor"][ok0])
    assert np.all(c1[~ok1] == -1) and np.all(c0[~ok0] == -1)
    assert ok1.sum() > 20000


def _proj_inputs(g):
    cams, classes, H, W = (int(v) for v in g["mask_shape"])
    mask = g["mask"] if g["mask"].size else synth.mask_planes(cams, classes, H, W, seed=int(g["mask_seed"]))
    return g["points"], g["lidar2img"], mask.reshape(cams, classes, H, W), H, W


@pytest.mark.parametrize("tag", ["projection_small", "projection_nusc"])
def test_projection(tag):
    g = load_golden(tag)
    pts, l2i, mask, H, W = _proj_inputs(g)
    p2d = O.prj_points_2d(pts, l2i, H, W)
    # the reference's K=4 matmul accumulation order is a BLAS detail; grid coords agree to a few ulp
    np.testing.assert_allclose(p2d, g["pts_2d"], rtol=2e-5, atol=2e-6)
    ids = O.points_in_mask(pts, mask, l2i)
    flips = np.any(ids != g["ids"].astype(np.int64), axis=(1, 2)).mean()
    assert flips <= 2e-3, f"id flip rate {flips}"       # texel-boundary flips only
    ids_sel, cam, fg, _ = O.cam_select(g["ids"].astype(np.int64))
    assert np.array_equal(cam, g["cam_sel"]) and np.array_equal(ids_sel, g["ids_sel"]) and np.array_equal(fg, g["fg"])


@pytest.mark.parametrize("tag", ["ccl_small", "ccl_mid", "ccl_batch"])
def test_ccl(tag):
    g = load_golden(tag)
    if tag == "ccl_batch":
        lab = O.connected_components(g["points"], g["batch_idx"], float(g["dist"]))
    else:
        lab = O.connected_components_single_batch(g["points"], float(g["dist"]))
    assert np.array_equal(lab, g["labels"])


@pytest.mark.parametrize("tag,norm,act,eps", [("mlp_ln_gelu", "ln", "gelu", 1e-3), ("mlp_head", "ln", "gelu", 1e-3),
                                              ("mlp_bn_relu", "bn", "relu", 1e-3)])
def test_mlp(tag, norm, act, eps):
    g = load_golden(tag)
    sd = {k.replace("__", "."): v for k, v in g.items() if k not in ("x", "y")}
    y = O.mlp_from_state_dict(g["x"], sd, norm, act, eps)
    np.testing.assert_allclose(y, g["y"], rtol=1e-4, atol=1e-5)


def test_neck():
    g = load_golden("neck")
    out, mask = O.voxel2point_neck(g["points"], g["coors"], g["voxel_feats"], g["voxel2point_inds"],
                                   [0.2, 0.2, 0.2], [-51.2, -51.2, -5.0, 51.2, 51.2, 3.0])
    assert np.array_equal(mask, g["mask"])
    assert np.array_equal(out, g["out"])


def test_ingroup():
    g = load_golden("ingroup")
    assert np.array_equal(O.ingroup_indices(g["group"]), g["inner"])


def test_vote_decode():
    g = load_golden("vote_decode")
    assert np.array_equal(O.decode_vote_targets(g["preds"]), g["offsets"])


# ---- query refinement: goldens written by `python tools/make_golden.py refine` -------------------
def test_box_decode_golden():
    """BasePointBBoxCoder.decode / FSF.decode_stage_bboxes (the reference's own code) vs the oracle."""
    g = load_golden("box_decode")
    got = O.decode_boxes(g["reg"], g["base"])
    np.testing.assert_allclose(got, g["rois"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(got[:, 1:], g["boxes"], rtol=1e-6, atol=1e-6)


def test_roi_extractor_golden():
    """The reference's DynamicPointROIExtractor (debug=True: its invariants were asserted when the golden was written) and
    DynamicPointPoolFunction around the oracle's pooling semantics: buffer protocol, valid-row filter, field slicing."""
    g = load_golden("roi_extractor")
    p, r, f = O.dynamic_point_pool(g["rois"][:, 1:8], g["points"], [1.0, 1.0, 1.0], 512, 50000)
    assert np.array_equal(p, g["inds"]) and np.array_equal(r, g["roi_inds"])
    np.testing.assert_array_equal(f[:, 3:6], g["local_xyz"])
    np.testing.assert_array_equal(f[:, 6:12], g["boundary_offset"])
    np.testing.assert_array_equal(f[:, 12], g["is_in_margin"])
    e = load_golden("roi_extractor_empty")   # nothing pooled: one fake row of -1 ids (dynamic_point_pool_op.py:36-40)
    assert e["inds"].tolist() == [-1] and e["roi_inds"].tolist() == [-1] and e["local_xyz"].shape == (1, 3)


def test_roi_align_golden():
    from oracle import fsf_oracle_models as OM
    g = load_golden("roi_align")
    new, mask = OM.align_roi_features(g["feats"], g["out_coors"], int(g["num_rois"]))
    assert np.array_equal(mask, g["mask"])
    np.testing.assert_array_equal(new, g["aligned"])


def test_group_cluster_golden():
    """The reference's own group_sample + ClusterAssigner (six class groups, both fallbacks hit) vs the oracle's restatement."""
    from fullysparsefusion_b200.fsf import NUSC
    from oracle import fsf_oracle_frame as OF
    g = load_golden("group_cluster")
    groups = [[NUSC["class_names"].index(n) for n in grp] for grp in NUSC["group_names"]]
    score, centers = OF.group_sample(g["logits"], g["points"], g["offsets"], groups, NUSC["score_thresh"])
    rows, inds, ctrs = [], [], []
    for gi in range(6):
        idx = np.flatnonzero(score[:, gi] > np.float32(NUSC["score_thresh"][gi]))
        if idx.size == 0:
            idx = np.zeros(1, np.int64)
        labels, keep = OF.cluster_assign_single(centers[idx, gi], NUSC["cluster_voxel_size"][gi], NUSC["point_cloud_range"],
                                                NUSC["connected_dist"][gi], NUSC["min_points"])
        rows.append(idx[keep])
        inds.append(np.stack([np.full(keep.size, gi), np.zeros(keep.size, np.int64), labels], 1))
        ctrs.append(centers[idx[keep], gi])
    assert [len(r) for r in rows] == g["counts"].tolist()
    assert np.array_equal(np.concatenate(rows), g["rows"])
    assert np.array_equal(np.concatenate(inds), g["cluster_inds"])
    np.testing.assert_allclose(np.concatenate(ctrs), g["center_preds"], rtol=1e-5, atol=1e-5)


# ---- more in-tree reference Python (`python tools/make_golden.py misc`) -----------------------------------------------
def test_pre_voxelize_golden():
    from oracle import fsf_oracle_frame as OF
    g = load_golden("pre_voxelize")
    data = dict(p=g["points"], l=g["logits"], v=g["votes"], f=g["feats"])
    vox, uniq, inv = OF.pre_voxelize(data, g["points"], (0.1, 0.1, 0.1), [-51.2, -51.2, -5.0, 51.2, 51.2, 3.0])
    assert len(uniq) == len(g["v_points"])          # same voxel set, same (torch.unique) order
    for k, want in (("p", "v_points"), ("l", "v_logits"), ("v", "v_votes"), ("f", "v_feats")):
        np.testing.assert_allclose(vox[k], g[want], rtol=1e-4, atol=1e-5)


def test_encode_preds_2d_golden():
    from oracle import fsf_oracle_frame as OF
    g = load_golden("encode_preds_2d")
    preds, enc = OF.encode_preds_2d(g["anno"], g["obj_coors"][:, 2], int(g["img_w"]), int(g["img_h"]), 10)
    np.testing.assert_array_equal(preds, g["preds_2d"])
    np.testing.assert_allclose(enc, g["enc"], rtol=1e-6, atol=1e-7)


def test_img_cross_attn_golden():
    """FSF.img_cross_attn up to the MLP input (the per-point, per-class 2D scores of the selected camera)."""
    from oracle import fsf_oracle_frame as OF
    g = load_golden("img_cross_attn")
    scores, ids, cam, fg, overlap = OF.img_scores(g["points"], g["mask"], g["lidar2img"], g["anno"])
    flips = (scores != g["scores"]).any(1).mean()
    assert flips <= 0.002, flips    # texel-boundary flips only (BLAS accumulation order of the K=4 projection)
    same = ~(scores != g["scores"]).any(1)
    np.testing.assert_array_equal(scores[same], g["scores"][same])


def test_fg_weights_golden():
    from oracle import fsf_oracle_frame as OF
    g = load_golden("fg_weights")
    np.testing.assert_allclose(OF.point_fg_weights(g["logits"]), g["weights"], rtol=1e-5, atol=1e-6)


def test_combine_golden():
    """combine_frustum_and_fsd's bookkeeping: concatenation order and the (batch, cls, id + fsd_begin_idx) rewrite of the
    LiDAR queries' coordinates — what fsf.FSF.stages()['combine'] does with torch.cat / torch.stack."""
    g = load_golden("combine")
    fc = g["fsd_coors"]
    fsd_re = np.stack([fc[:, 1], fc[:, 0], fc[:, 2] + 1000], 1)
    assert np.array_equal(np.concatenate([g["fr_coors"], fsd_re]), g["obj_coors"])
    assert np.array_equal(np.concatenate([g["fr_centers"], g["fsd_centers"]]), g["obj_centers"])
    assert np.array_equal(np.concatenate([g["fr_cls"], g["fsd_cls"]]), g["obj_cls"])
    assert np.array_equal(np.concatenate([g["fr_reg"], g["fsd_reg"]]), g["obj_reg"])
    assert (g["preds_2d"][len(g["fr_coors"]):] == 0).all()


def _v2_head_from_golden():
    import torch
    from fullysparsefusion_b200 import fsf as FSFM
    g = load_golden("cluster_head_v2")
    names = ["car", "truck", "bus"]
    head = FSFM.SparseClusterHeadV2(num_classes=3, in_channel=24, shared_mlp_dims=[32, 32], tasks=[dict(class_names=names)],
                                    common_attrs=dict(center=(3, 2, 16), dim=(3, 2, 16), rot=(2, 2, 16), vel=(2, 2, 16)),
                                    num_cls_layer=2, cls_hidden_dim=16, separate_head=dict(norm_cfg=dict(type="LN"), act="gelu"),
                                    norm_cfg=dict(type="LN"), act="relu")
    sd = {k.replace("__", "."): torch.from_numpy(v) for k, v in g.items() if k not in ("x", "cls", "reg")}
    return g, head, sd


def test_cluster_head_state_dict_compat():
    """A state dict saved by the REFERENCE's SparseClusterHeadV2 loads strictly into fsf.SparseClusterHeadV2 (same keys and
    shapes: the checkpoint layout the stock configs produce), and the oracle reproduces the reference's outputs from it."""
    g, head, sd = _v2_head_from_golden()
    missing, unexpected = head.load_state_dict(sd, strict=True)
    assert not missing and not unexpected
    sdn = {k: v.numpy() for k, v in sd.items()}
    h = O.mlp_from_state_dict(g["x"], {k[len("shared_mlp."):]: v for k, v in sdn.items() if k.startswith("shared_mlp.")}, "ln", "relu", 1e-5)
    outs = {a: O.mlp_from_state_dict(h, {k[len(f"task_heads.0.{a}."):]: v for k, v in sdn.items() if k.startswith(f"task_heads.0.{a}.")},
                                     "ln", "gelu", 1e-5) for a in ("center", "dim", "rot", "vel", "score")}
    np.testing.assert_allclose(outs["score"], g["cls"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(np.concatenate([outs["center"], outs["dim"], outs["rot"], outs["vel"]], 1), g["reg"], rtol=1e-4, atol=1e-5)


def _seg_head_from_golden():
    import torch
    from fullysparsefusion_b200 import modules as M

    g = load_golden("vote_seg_head")
    head = M.VoteSegHead(in_channel=24, num_classes=4, hidden_dims=[16, 16], dropout_ratio=0.0, norm_cfg=dict(type="naiveSyncBN1d"),
                         act_cfg=dict(type="ReLU"))
    sd = {k.replace("__", "."): torch.from_numpy(np.asarray(v)) for k, v in g.items() if k not in ("x", "logits", "votes", "num_classes")}
    return g, head, sd


def test_vote_seg_head_state_dict_compat():
    """A state dict saved by the REFERENCE's VoteSegHead (stock loss config: background class appended) loads strictly into
    modules.VoteSegHead, and the oracle reproduces the reference's logits and votes from it."""
    from oracle import fsf_oracle_models as OM

    g, head, sd = _seg_head_from_golden()
    missing, unexpected = head.load_state_dict(sd, strict=True)
    assert not missing and not unexpected
    assert int(g["num_classes"]) == head.num_classes == 5 and g["logits"].shape == (300, 5) and g["votes"].shape == (300, 15)
    logits, votes = OM.vote_seg_head(g["x"], {k: v.numpy() for k, v in sd.items()})
    np.testing.assert_allclose(logits, g["logits"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(votes, g["votes"], rtol=1e-4, atol=1e-5)
