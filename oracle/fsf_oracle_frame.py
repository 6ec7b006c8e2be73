"""CPU oracle of the FSF frame-level glue (query generation) — TEST INFRASTRUCTURE ONLY.

numpy restatements of the in-tree reference functions between the learned blocks; every function cites
the reference lines it follows.  extract_fg_pts / double_overlap_pts / get_sir_coors /
get_cluster_delta_weighted are pinned by tests/golden/frustum_pool.npz (outputs of the reference's own
functions); the rest follow the in-tree Python literally."""
from __future__ import annotations

import numpy as np

from . import fsf_oracle as O

F32 = np.float32


def softmax32(logits):
    """torch softmax(dim=1) in fp32: exp(x - max) / sum."""
    l = np.asarray(logits, F32)
    e = np.exp((l - l.max(1, keepdims=True)).astype(F32)).astype(F32)
    s = np.zeros(len(l), F32)
    for j in range(l.shape[1]):
        s = (s + e[:, j]).astype(F32)
    return (e / s[:, None]).astype(F32)


def point_fg_weights(seg_logits):
    """FSF.get_point_fg_weights (FSF.py:345-355)."""
    return (F32(1) - softmax32(seg_logits)[:, -1]).astype(F32)


def img_scores(points_noaug, mask, lidar2img, anno, col=4):
    """FSF.img_cross_attn up to the MLP input on nuScenes (FSF.py:694-728, 506-552):
    ids [N,6,10] → camera with the largest id sum → mask_anno[id-1][score] per class (0 where id == 0)."""
    ids = O.points_in_mask(points_noaug, mask, lidar2img)
    ids_sel, cam, fg, overlap = O.cam_select(ids)
    scores = np.where(ids_sel > 0, np.asarray(anno, F32)[np.clip(ids_sel - 1, 0, len(anno) - 1), col], F32(0)).astype(F32)
    return scores, ids, cam, fg, overlap


def frustum_rows(ids):
    """extract_fg_pts + double_overlap_pts + get_sir_coors (FSF.py:260-308, 357-365) for one sample:
    returns (rows_point [R] — source point per output row, obj_id [R])."""
    n = ids.shape[0]
    flat = ids.reshape(n, -1)
    fg = np.flatnonzero(flat.sum(1) > 0)
    flat = flat[fg]
    overlaps = (flat > 0).sum(1)
    rows = [fg]
    obj = [flat.max(1)] if len(fg) else [np.zeros(0, flat.dtype)]
    for k in range(2, int(overlaps.max()) + 1 if len(fg) else 0):
        m = overlaps == k
        if m.sum() == 0:
            continue
        rows.append(np.tile(fg[m], k - 1))                       # .repeat(overlap_num - 1, 1)
        srt = -np.sort(-flat[m], axis=1)[:, :k]                  # topk values, descending
        for pad in range(1, k):
            obj.append(srt[:, pad])
    return np.concatenate(rows), np.concatenate(obj)


def cluster_delta_weighted(points_xyz, sir_coors, weights):
    """get_cluster_delta_weighted (FSF.py:313-329) → (f_cluster, centre, coors, inv)."""
    w = np.maximum(np.asarray(weights, F32), F32(1e-5))[:, None]
    feat = np.concatenate([(np.asarray(points_xyz, F32)[:, :3] * w).astype(F32), w], 1)
    mean, coors, inv = O.scatter_v2(feat, sir_coors, "avg")
    center = (mean[:, :3] / mean[:, 3:4]).astype(F32)
    return (np.asarray(points_xyz, F32)[:, :3] - center[inv]).astype(F32), center, coors, inv


def encode_preds_2d(anno, obj_ids, img_w, img_h, num_classes):
    """get_single_cls_preds_2d + encode_preds_2d(encode_single_cls=True) (FSF.py:449-504)."""
    anno = np.asarray(anno, F32)
    k = len(obj_ids)
    preds = np.zeros((k, anno.shape[1]), F32)
    idx = np.asarray(obj_ids, np.int64) - 1
    ok = idx >= 0
    preds[ok] = anno[idx[ok]]
    preds[~ok, 5] = num_classes
    box = preds[:, :4].copy()
    box[:, 0::2] = (box[:, 0::2] / F32(img_w)).astype(F32)
    box[:, 1::2] = (box[:, 1::2] / F32(img_h)).astype(F32)
    onehot = np.eye(num_classes + 1, dtype=F32)[preds[:, 5].astype(np.int64)]
    return preds, np.concatenate([box, preds[:, 4:5], onehot], 1)


def pre_voxelize(data: dict, points, voxel_size, pc_range):
    """SingleStageFSD.pre_voxelize (single_stage_fsd.py:585-605), one sample."""
    c = O.voxelize(points, voxel_size, pc_range, floor_mode=1, grid=[1 << 20] * 3).astype(np.int64)
    coors = np.concatenate([np.zeros((len(c), 1), np.int64), c], 1)
    uniq, inv, _ = O.unique_rows(coors)
    return {k: O.scatter_mean(v, inv) for k, v in data.items()}, uniq, inv


def group_sample(seg_logits, seg_points, offsets, groups, thresholds):
    """SingleStageFSD.group_sample (single_stage_fsd.py:802-865) + get_offset_weight('max', :868-874):
    (group_score [n,G], centres [n,G,3]); fg_mask_g = group_score[:, g] > thr."""
    l = np.asarray(seg_logits, F32)
    p = softmax32(l)
    n, c1 = l.shape
    off = np.asarray(offsets, F32).reshape(n, c1, 3)
    score = np.zeros((n, len(groups)), F32)
    centers = np.zeros((n, len(groups), 3), F32)
    for g, idx in enumerate(groups):
        s = np.zeros(n, F32)
        for c in idx:
            s = (s + p[:, c]).astype(F32)
        score[:, g] = s
        lg = l[:, idx]
        w = (np.abs((lg - lg.max(1, keepdims=True)).astype(F32)) < F32(1e-6)).astype(F32)
        w = (w / w.sum(1, keepdims=True).astype(F32)).astype(F32)
        o = np.zeros((n, 3), F32)
        for t, c in enumerate(idx):
            o = (o + (off[:, c, :] * w[:, t:t + 1]).astype(F32)).astype(F32)
        centers[:, g] = (np.asarray(seg_points, F32)[:, :3] + o).astype(F32)
    del thresholds
    return score, centers


def cluster_assign_single(centers, voxel_size, pc_range, dist, min_points):
    """ClusterAssigner.forward_single_class, inference / single-batch variant (single_stage_fsd.py:936-982):
    returns (cluster id per KEPT point, kept indices)."""
    c = O.voxelize(centers, voxel_size, pc_range, floor_mode=1, grid=[1 << 30] * 3)
    lo = np.asarray(pc_range[:3], F32)
    vs = np.asarray(voxel_size, F32)
    # raw (unchecked) coordinates in x,y,z order: recompute without range rejection
    a = (np.asarray(centers, F32)[:, :3] - lo[None]).astype(F32)
    mod = np.fmod(a, vs[None]).astype(F32)
    div = ((a - mod).astype(F32) / vs[None]).astype(F32)
    adj = (mod != 0) & (mod < 0)
    div = np.where(adj, (div - F32(1)).astype(F32), div)
    fl = np.floor(div)
    fl = np.where((div - fl).astype(F32) > F32(0.5), fl + F32(1), fl)
    cx = np.where(div != 0, fl, F32(0)).astype(np.int64)
    del c
    coors = np.concatenate([np.zeros((len(cx), 1), np.int64), cx], 1)
    _, inv, cnt = O.unique_rows(coors)
    valid = cnt[inv] >= min_points
    if not valid.any():
        valid = ~valid
    keep = np.flatnonzero(valid)
    sampled, _, inv2 = O.scatter_v2(np.asarray(centers, F32)[keep], coors[keep], "avg")
    labels = O.connected_components_single_batch(sampled, dist)
    return labels[inv2], keep
