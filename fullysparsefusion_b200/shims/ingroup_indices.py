"""ingroup_indices extension module — projects/mmdet3d_plugin/ops/sst_ops.py:239,246-248:
    ingroup_indices.forward(group_inds, out_inds) -> None      (out_inds caller-allocated, filled in place)
Each element receives its rank inside its group; this implementation is the stable rank (original order),
equal to the reference's slow oracle get_inner_win_inds_slow (middle_encoders/sst_input_layer.py:200-208)."""
from __future__ import annotations

import torch

from .. import ops


def forward(group_inds: torch.Tensor, out_inds: torch.Tensor) -> None:
    assert group_inds.dim() == 1 and out_inds.shape == group_inds.shape
    assert out_inds.dtype == torch.int64 and group_inds.dtype == torch.int64
    out_inds.copy_(ops.ingroup_indices(group_inds))
