"""mmdet3d.ops names the plugin imports: Voxelization (dynamic mode) and furthest_point_sample (single_stage_fsd.py:13), and the
`spconv` module object (ops/sst_ops.py:5).

    Voxelization(voxel_size, point_cloud_range, max_num_points=-1, max_voxels=(-1,-1))(points) -> coors[N,3] (z,y,x)
Only the dynamic mode the FSF configs use (max_num_points=-1, FSF_nuScenes_config.py:36-41) exists here."""
from __future__ import annotations

import torch
from torch import nn

from .. import ops


class Voxelization(nn.Module):
    def __init__(self, voxel_size, point_cloud_range, max_num_points=-1, max_voxels=(-1, -1), deterministic=True):
        super().__init__()
        if max_num_points != -1:
            raise NotImplementedError("only dynamic voxelization (max_num_points=-1) is on the FSF path")
        self.voxel_size = list(voxel_size)
        self.point_cloud_range = list(point_cloud_range)
        self.max_num_points = max_num_points
        self.max_voxels = max_voxels
        self.grid_size = torch.tensor(ops.grid_shape(point_cloud_range, voxel_size))

    def forward(self, points: torch.Tensor) -> torch.Tensor:
        return ops.voxelize(points, self.voxel_size, self.point_cloud_range, floor_mode=0)

    def __repr__(self):
        return (f"{self.__class__.__name__}(voxel_size={self.voxel_size}, point_cloud_range={self.point_cloud_range}, "
                f"max_num_points={self.max_num_points}, max_voxels={self.max_voxels})")


def furthest_point_sample(points: torch.Tensor, num_points: int) -> torch.Tensor:
    """mmdet3d.ops.furthest_point_sample(points [B,N,3], M) -> idx [B,M] int32: imported at single_stage_fsd.py:13 and used only
    by the `fps` helper (:25-29), which no stock FSF config reaches (ClusterAssigner is built without it).  Provided so the import
    succeeds and the helper works: the iterative farthest-point rule (start at point 0, repeatedly take the point farthest from
    the chosen set), evaluated with torch ops on the tensor's own device — off the hot path."""
    assert points.dim() == 3 and points.size(2) >= 3
    b, n, _ = points.shape
    idx = torch.zeros((b, num_points), dtype=torch.int32, device=points.device)
    dist = torch.full((b, n), float("inf"), device=points.device)
    far = torch.zeros(b, dtype=torch.long, device=points.device)
    ar = torch.arange(b, device=points.device)
    for i in range(num_points):
        idx[:, i] = far.to(torch.int32)
        d = ((points[:, :, :3] - points[ar, far, :3][:, None]) ** 2).sum(-1)
        dist = torch.minimum(dist, d)
        far = dist.argmax(1)
    return idx


class _SpconvModule:
    """Stand-in for the `mmdet3d.ops.spconv` module object sst_ops.py imports at load time (:5).  The names the plugin touches
    at import time are classes it subclasses or references lazily; the FSF forward path never calls spconv through sst_ops (the
    sparse U-Net is the registry type `SimpleSparseUNet`, served by modules.SimpleSparseUNet).  Attribute access therefore hands
    back placeholders that raise on use, naming the B200 replacement."""

    class SparseConvTensor:   # sst_ops.py constructs it only inside SST-specific helpers that FSF does not call
        def __init__(self, features, indices, spatial_shape, batch_size):
            self.features, self.indices, self.spatial_shape, self.batch_size = features, indices, spatial_shape, batch_size

    def __getattr__(self, name):
        def _unavailable(*a, **k):
            raise NotImplementedError(f"mmdet3d.ops.spconv.{name}: the B200 path serves sparse convolutions through the registry type "
                                      "SimpleSparseUNet (fullysparsefusion_b200.modules) and fsfb_gather_gemm")
        _unavailable.__name__ = name
        return _unavailable


spconv = _SpconvModule()
