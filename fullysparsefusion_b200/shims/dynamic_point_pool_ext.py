"""dynamic_point_pool_ext extension module — projects/mmdet3d_plugin/ops/dynamic_point_pool_op.py:5,27-32:
    dynamic_point_pool_ext.forward(rois, pts, extra_wlh, max_inbox_point, out_pts_idx, out_roi_idx, out_pts_feats) -> None
Outputs are caller-allocated and prefilled (-1 ids, zero features, 50000 rows upstream); valid rows are those with
out_pts_idx >= 0 (:34).  This implementation fills rows [0, count) in (roi, point) order and leaves the rest as given."""
from __future__ import annotations

from typing import Sequence

import torch

from .. import ops


def forward(rois: torch.Tensor, pts: torch.Tensor, extra_wlh: Sequence[float], max_inbox_point: int,
            out_pts_idx: torch.Tensor, out_roi_idx: torch.Tensor, out_pts_feats: torch.Tensor) -> None:
    assert rois.dim() == 2 and rois.size(1) == 7, "rois: [K,7] (x,y,z,w,l,h,rz)"
    ops.dynamic_point_pool(rois.float(), pts[:, :3].float(), list(extra_wlh), int(max_inbox_point), out_pts_idx, out_roi_idx,
                           out_pts_feats)
