"""CPU port of the reference's forward sequence in stock torch ops — TEST/BENCH INFRASTRUCTURE ONLY.

This is the "reference's CPU torch_scatter / SIR path" of BASELINE.json's north_star, restated with
the ATen CPU operators the reference's Python calls (or, for torch_scatter 2.0.2 — absent here —
their nearest ATen equivalents scatter_reduce_/index_add_).  It is multi-threaded through
torch.set_num_threads and is what bench.py times as `cpu_baseline` / `--impl reference`
(kind "port": the reference itself cannot be installed — no mmcv/mmdet3d/spconv/torch_scatter).
Only tests/ and bench.py may import this module.  Parity pinning: see oracle/fsf_oracle.py.

Stage names match fullysparsefusion_b200/frame.py.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Tuple

import torch
import torch.nn.functional as F

NUSC_RANGE = (-51.2, -51.2, -5.0, 51.2, 51.2, 3.0)


def _voxel_coors(points, voxel, rng):
    """torch.div(p - min, vs, rounding_mode='floor') in zyx order, batch-padded
    (single_stage_fsd.py:222-225, :591-593)."""
    lo = torch.tensor(rng[:3], dtype=torch.float32)
    vs = torch.tensor(voxel, dtype=torch.float32)
    c = torch.div(points[:, :3] - lo[None], vs[None], rounding_mode="floor").long()[:, [2, 1, 0]]
    return F.pad(c, (1, 0), value=0)


def scatter_mean(feat, inv, m):
    """torch_scatter.scatter(reduce='mean') (sst_ops.py:170)."""
    s = torch.zeros((m, feat.size(1)), dtype=feat.dtype).index_add_(0, inv, feat)
    cnt = torch.bincount(inv, minlength=m).clamp(min=1).to(feat.dtype)
    return s / cnt[:, None]


def scatter_max(feat, inv, m):
    """torch_scatter.scatter_max values (sst_ops.py:168; argmax is discarded by scatter_v2)."""
    out = torch.full((m, feat.size(1)), float("-inf"), dtype=feat.dtype)
    return out.scatter_reduce_(0, inv[:, None].expand_as(feat), feat, reduce="amax", include_self=True)


def prj_points_2d(points, lidar2img, img_h, img_w):
    """FSF.prj_points_2d, op for op (FSF.py:169-200)."""
    pts_4d = torch.cat([points[:, :3], points.new_ones((points.size(0), 1))], dim=-1)
    pts_2d = pts_4d @ lidar2img.permute(0, 2, 1)
    depth_valid = pts_2d[..., 2] > 1e-3
    pts_2d[..., 2] = torch.clamp(pts_2d[..., 2], min=1e-5, max=1e5)
    pts_2d[..., 0] /= pts_2d[..., 2]
    pts_2d[..., 1] /= pts_2d[..., 2]
    pts_2d[..., 0] /= img_w
    pts_2d[..., 1] /= img_h
    pts_2d = pts_2d[..., :2]
    pts_2d = (pts_2d - 0.5) * 2
    valid = depth_valid & (pts_2d[..., 0] > -1) & (pts_2d[..., 0] < 1) & (pts_2d[..., 1] > -1) & (pts_2d[..., 1] < 1)
    pts_2d[~valid] = -2.0
    return pts_2d


def points_in_mask(points, mask_data, lidar2img):
    """FSF.points_in_mask (FSF.py:202-226): float cast of the planes, one grid_sample per camera."""
    cams, classes, H, W = mask_data.shape
    pts_2d = prj_points_2d(points, lidar2img, H, W)
    mask_f = mask_data.float()
    out = []
    for cam in range(cams):
        s = F.grid_sample(mask_f[cam][None], pts_2d[cam][None, None], mode="nearest", align_corners=False)
        out.append(s[0, :, 0, :].permute(1, 0))
    return torch.stack(out, 1).long()


def build_stages(host: Dict[str, torch.Tensor], seed: int = 0) -> Tuple[List[Tuple[str, Callable[[], None]]], dict]:
    points, mask, l2i = host["points"], host["mask"], host["lidar2img"]
    n = points.size(0)
    g = torch.Generator().manual_seed(seed)
    pt_feats = torch.randn(n, 64, generator=g)
    seg_logits, vote_preds, seg_feats = torch.randn(n, 11, generator=g), torch.randn(n, 33, generator=g), torch.randn(n, 131, generator=g)
    st: dict = {}

    def voxelize():
        st["coors"] = _voxel_coors(points, (0.2, 0.2, 0.2), NUSC_RANGE)

    def rank():
        st["voxel_coors"], st["inv"] = torch.unique(st["coors"], return_inverse=True, dim=0)

    def csr():
        pass  # the reference has no rulebook for scatters: every call re-walks the index

    def vfe_scatter():
        m = st["voxel_coors"].size(0)
        st["voxel_mean"] = scatter_mean(points[:, :5], st["inv"], m)
        st["vfe0"] = scatter_max(pt_feats, st["inv"], m)
        st["vfe1"] = scatter_max(pt_feats, st["inv"], m)

    def neck():
        st["pt_voxel_feats"] = torch.cat([st["vfe0"][st["inv"]], st["vfe1"][st["inv"]]], 1)

    def project():
        ids = points_in_mask(points[:, 5:8], mask, l2i)
        cam = ids.sum(-1).max(-1)[1]
        st["cam_sel"] = cam
        st["ids_sel"] = ids[torch.arange(n), cam]
        st["fg"] = ids.sum((-2, -1)) > 0

    def pre_voxelize():
        coors = _voxel_coors(points, (0.1, 0.1, 0.1), NUSC_RANGE)
        uniq, inv = torch.unique(coors, return_inverse=True, dim=0)
        m = uniq.size(0)
        st["pre_coors"] = uniq
        st["pre_points"] = scatter_mean(points[:, :5], inv, m)
        st["pre_logits"] = scatter_mean(seg_logits, inv, m)
        st["pre_votes"] = scatter_mean(vote_preds, inv, m)
        st["pre_feats"] = scatter_mean(seg_feats, inv, m)
        st["pre_centers"] = scatter_mean(vote_preds, inv, m)

    return [("voxelize", voxelize), ("rank", rank), ("csr", csr), ("vfe_scatter", vfe_scatter), ("neck", neck),
            ("project", project), ("pre_voxelize", pre_voxelize)], st
