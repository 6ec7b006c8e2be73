"""CPU oracle of the model-level blocks on the forward hot path — TEST INFRASTRUCTURE ONLY.

numpy restatements, written independently of fullysparsefusion_b200/modules.py, of the blocks the stock
configs name.  Parity status:
  * VoteSegHead, build_mlp stacks, Voxel2PointScatterNeck, SIR (the block loop): in-tree reference code,
    pinned through tests/golden (see fsf_oracle.py).
  * DynamicScatterVFE, SIRLayer/DynamicVFELayer, SimpleSparseUNet/SparseBasicBlock: NOT in the reference
    tree (un-vendored mmdet3d fork) → "parity unpinned": these follow the published
    SST/FSD + mmdet3d SparseUNet implementations; the configs pin their channel arithmetic
    (FSF_nuScenes_config.py:42-70, 113-124).
Weights come in as a state_dict of numpy arrays with the module's parameter names.
"""
from __future__ import annotations

import numpy as np

from . import fsf_oracle as O

F32 = np.float32


def _bn_layer(x, sd, prefix, eps, act):
    return O.mlp_layer(x, sd[prefix + "0.weight"], None, norm="bn", norm_w=sd[prefix + "1.weight"], norm_b=sd[prefix + "1.bias"],
                       mean=sd[prefix + "1.running_mean"], var=sd[prefix + "1.running_var"], eps=eps, act=act)


def dynamic_scatter_vfe(features, coors, sd, voxel_size, point_cloud_range, eps=1e-3):
    """DynamicScatterVFE(with_cluster_center, with_voxel_center, mode='max') forward → (voxel_feats, voxel_coors, inv)."""
    features = np.asarray(features, F32)
    uniq, inv, _ = O.unique_rows(coors)
    mean = O.scatter_mean(features, inv)
    f_cluster = (features[:, :3] - mean[inv][:, :3]).astype(F32)
    vs = np.asarray(voxel_size, F32)
    off = np.asarray([float(voxel_size[a]) / 2 + float(point_cloud_range[a]) for a in range(3)]).astype(F32)
    centre = ((np.asarray(coors)[:, [3, 2, 1]].astype(F32) * vs[None]).astype(F32) + off[None]).astype(F32)
    f_center = (features[:, :3] - centre).astype(F32)
    x = np.concatenate([features, f_cluster, f_center], 1)
    n_layers = len({k.split(".")[1] for k in sd if k.startswith("vfe_layers.")})
    vf = None
    for i in range(n_layers):
        pf = _bn_layer(x, sd, f"vfe_layers.{i}.", eps, "relu")
        vf = O.scatter_max(pf, inv)[0]
        if i != n_layers - 1:
            x = np.concatenate([pf, vf[inv]], 1)
    return vf, uniq, inv


def _sub(sd, prefix):
    return {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}


def sir_layer(features, inv, f_cluster, sd, xyz_normalizer, rel_dist_scaler=10.0, eps=1e-3, act="gelu", m=None):
    """SIRLayer forward (LN + act): returns (point_feats_with_shortcut, cluster_feats)."""
    features = np.asarray(features, F32)
    x = features.copy()
    x[:, :3] = (features[:, :3] / np.asarray(xyz_normalizer, F32)[None]).astype(F32)
    gate = O.mlp_from_state_dict((np.asarray(f_cluster, F32) / F32(rel_dist_scaler)).astype(F32), _sub(sd, "rel_mlp."), "ln", act, eps)
    x = (x * gate).astype(F32)
    ori = x
    n_layers = len({k.split(".")[1] for k in sd if k.startswith("vfe_layers.")})
    clusters = []
    pf = None
    for i in range(n_layers):
        p = f"vfe_layers.{i}."
        pf = O.mlp_layer(x, sd[p + "linear.weight"], None, norm="ln", norm_w=sd[p + "norm.weight"], norm_b=sd[p + "norm.bias"],
                         eps=eps, act=act)
        c = O.scatter_max(pf, inv, m)[0]
        clusters.append(c)
        if i != n_layers - 1:
            x = np.concatenate([pf, c[inv]], 1)
    if pf.shape == ori.shape:
        pf = (pf + ori).astype(F32)
    return pf, np.concatenate(clusters, 1)


def sir(points, features, coors, f_cluster, sd, num_blocks, xyz_normalizer, eps=1e-3, act="gelu"):
    """SIR.forward (models/backbones/sir.py:65-85) → (out_feats, cluster_feats [K, 2*C*blocks], out_coors)."""
    uniq, inv, _ = O.unique_rows(coors)
    out = np.asarray(features, F32)
    cl = []
    for i in range(num_blocks):
        in_feats = np.concatenate([np.asarray(points, F32), out], 1)
        out, c = sir_layer(in_feats, inv, f_cluster, _sub(sd, f"block_list.{i}."), xyz_normalizer, 10.0, eps, act, len(uniq))
        cl.append(c)
    return out, np.concatenate(cl, 1), uniq


def _conv_module(x, nbr, sd, prefix, eps, act="relu", residual=None):
    scale = sd[prefix + "bn.weight"] / np.sqrt(sd[prefix + "bn.running_var"] + F32(eps))
    shift = sd[prefix + "bn.bias"] - sd[prefix + "bn.running_mean"] * scale
    return O.gather_gemm(x, sd[prefix + "weight"], nbr, norm="affine", norm_w=scale, norm_b=shift, residual=residual, act=act)


def simple_sparse_unet(voxel_feats, voxel_coors, sd, sparse_shape, encoder_channels, encoder_paddings, decoder_channels,
                       batch_size=1, eps=1e-3):
    """SimpleSparseUNet.forward: conv_input, encoder stages (strided SparseConv3d first in stages > 0), decoder levels of
    lateral SparseBasicBlock + concat + merge + reduce_channel add + SparseInverseConv3d / final SubMConv3d."""
    def trip(p):
        return [int(p)] * 3 if isinstance(p, int) else [int(v) for v in p]

    coors = np.asarray(voxel_coors, np.int64)
    shape = [batch_size] + list(sparse_shape)
    levels = [dict(coors=coors, shape=shape)]
    rb = {"subm1": O.conv_rulebook(coors, coors, shape, (3, 3, 3), (1, 1, 1), (1, 1, 1))}
    for i in range(1, len(encoder_channels)):
        pad = trip(tuple(encoder_paddings[i])[0])
        prev = levels[-1]
        oshape = [batch_size] + [(prev["shape"][1 + a] + 2 * pad[a] - 3) // 2 + 1 for a in range(3)]
        oc = O.conv_out_coors(prev["coors"], oshape, (3, 3, 3), (2, 2, 2), pad)
        rb[f"spconv{i + 1}"] = O.conv_rulebook(oc, prev["coors"], prev["shape"], (3, 3, 3), (2, 2, 2), pad)
        rb[f"spconv{i + 1}_inv"] = O.conv_rulebook(prev["coors"], oc, oshape, (3, 3, 3), (2, 2, 2), pad, transposed=True)
        rb[f"subm{i + 1}"] = O.conv_rulebook(oc, oc, oshape, (3, 3, 3), (1, 1, 1), (1, 1, 1))
        levels.append(dict(coors=oc, shape=oshape))
    x = _conv_module(np.asarray(voxel_feats, F32), rb["subm1"], sd, "conv_input.", eps)
    enc = []
    for i, blocks in enumerate(encoder_channels):
        for j in range(len(blocks)):
            key = f"spconv{i + 1}" if (i != 0 and j == 0) else f"subm{i + 1}"
            x = _conv_module(x, rb[key], sd, f"encoder_layers.{i}.{j}.", eps)
        enc.append(x)
    x = enc[-1]
    for lvl in range(len(encoder_channels), 0, -1):
        lat_in = enc[lvl - 1]
        nb = rb[f"subm{lvl}"]
        h = _conv_module(lat_in, nb, sd, f"lateral_layer{lvl}.conv1.", eps)
        lat = _conv_module(h, nb, sd, f"lateral_layer{lvl}.conv2.", eps, residual=lat_in)
        cat = np.concatenate([x, lat], 1)
        merged = _conv_module(cat, nb, sd, f"merge_layer{lvl}.", eps)
        cm = merged.shape[1]
        reduced = cat.reshape(cat.shape[0], cm, -1).sum(2, dtype=F32)
        x = (merged + reduced).astype(F32)
        x = _conv_module(x, rb[f"spconv{lvl}_inv"] if lvl != 1 else rb["subm1"], sd, f"upsample_layer{lvl}.", eps)
    return x, rb, levels


def vote_seg_head(feats, sd, eps=1e-5):
    """VoteSegHead.forward (segmentation_head.py:89-104): pre_seg_conv (Linear→BN→ReLU)*, conv_seg, voting."""
    x = O.mlp_from_state_dict(np.asarray(feats, F32), _sub(sd, "pre_seg_conv."), "bn", "relu", eps)
    logits = O.mlp_layer(x, sd["conv_seg.weight"], sd["conv_seg.bias"])
    votes = O.mlp_layer(x, sd["voting.weight"], sd["voting.bias"])
    return logits, votes


def align_roi_features(feats, out_coors, num_rois):
    """FullySparseBboxHead.get_nonempty_roi_mask + align_roi_feature_and_rois (fsd_bbox_head.py:152-197) — pinned by
    tests/golden/roi_align.npz (the reference's own methods)."""
    feats = np.asarray(feats, F32)
    ids = np.asarray(out_coors, np.int64)
    new = np.zeros((num_rois, feats.shape[1]), F32)
    mask = np.zeros(num_rois, bool)
    new[ids[ids >= 0]] = feats[ids >= 0]
    mask[ids[ids >= 0]] = True
    return new, mask


def fully_sparse_bbox_head(pts_xyz, pts_features, local_xyz, boundary_offset, is_in_margin, roi_inds, rois, sd, num_blocks,
                           xyz_normalizer=(20, 20, 4), eps=1e-3, act="gelu", geo_input=True):
    """FullySparseBboxHead.forward (models/roi_heads/bbox_heads/fsd_bbox_head.py:95-151) with DynamicClusterVFE restated as the
    SIRLayer above (PARITY UNPINNED: un-vendored registry type).  Returns (roi feats [K, C], nonempty mask [K])."""
    pts_xyz = np.asarray(pts_xyz, F32)
    rois = np.asarray(rois, F32)
    roi_inds = np.asarray(roi_inds, np.int64)
    rel_xyz = (pts_xyz[:, :3] - rois[:, 1:4][roi_inds]).astype(F32)
    f_cluster = np.concatenate([local_xyz, boundary_offset, np.asarray(is_in_margin, F32)[:, None], rel_xyz], 1).astype(F32)
    uniq, inv, _ = O.unique_rows(roi_inds[:, None])
    out = np.asarray(pts_features, F32)
    cl = []
    for i in range(num_blocks):
        parts = [pts_xyz, out] + ([(f_cluster / F32(10.0)).astype(F32)] if geo_input else [])
        out, c = sir_layer(np.concatenate(parts, 1), inv, f_cluster, _sub(sd, f"block_list.{i}."), xyz_normalizer, 10.0, eps, act, len(uniq))
        cl.append(c)
    feats = np.concatenate(cl, 1)
    return align_roi_features(feats, uniq[:, 0], rois.shape[0])
