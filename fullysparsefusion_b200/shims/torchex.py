"""torchex.connected_components and the CCL helpers of single_stage_fsd.py.

    torchex.connected_components(points[m,3] f32, batch_idx[m] i32, dist, max_neighbor=100, mode=2, check=False)
        -> labels[m] int32                                        (call site single_stage_fsd.py:41)
    find_connected_componets(points, batch_idx, dist)              (single_stage_fsd.py:45-67)
    find_connected_componets_single_batch(points, batch_idx, dist) (single_stage_fsd.py:69-82)
Labels are contiguous from 0 (the reference asserts len(unique) == max+1, :42) and numbered exactly as
scipy.sparse.csgraph.connected_components numbers them on the reference's dense adjacency."""
from __future__ import annotations

import torch

from .. import ops


def connected_components(points: torch.Tensor, batch_idx: torch.Tensor, dist: float, max_neighbor: int = 100,
                         mode: int = 2, check: bool = False) -> torch.Tensor:
    del max_neighbor, mode, check  # neighbour caps of the TorchEx kernel do not apply: every pair is tested
    assert len(points) > 0
    return ops.connected_components(points, batch_idx, dist)


def find_connected_componets(points: torch.Tensor, batch_idx: torch.Tensor, dist: float) -> torch.Tensor:
    return ops.connected_components(points, batch_idx, dist).to(batch_idx.dtype)


def find_connected_componets_single_batch(points: torch.Tensor, batch_idx: torch.Tensor, dist: float) -> torch.Tensor:
    del batch_idx  # ignored by the reference as well (:69-82)
    return ops.connected_components(points, None, dist)
