"""mmdet3d.ops names the plugin imports (single_stage_fsd.py:13): Voxelization (dynamic mode).

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
