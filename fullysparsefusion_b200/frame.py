"""One frame of the FSF sparse forward path, as a list of named stages over device tensors.

Each stage is a closure over the C-ABI ops (fullysparsefusion_b200.ops); bench.py times every
stage with CUDA events and the tests compare stage outputs with the CPU oracle.  The stage order
follows the reference's inference call stack:

  voxelize      VoteSegmentor.voxelize                      single_stage_fsd.py:206-226
  rank          torch.unique(coors, dim=0) in scatter_v2    sst_ops.py:156
  csr           (the scatter rulebook shared by all reductions over one ranking)
  vfe_scatter   DynamicScatterVFE's mean + 2x max scatters  FSF_nuScenes_config.py:42-52
  neck          Voxel2PointScatterNeck gather               voxel2point_neck.py:42-67
  project       FSF.frustum_gather + camera selection       FSF.py:202-258, 714-718
  pre_voxelize  SingleStageFSD.pre_voxelize (0.1 m means)   single_stage_fsd.py:585-605

No oracle import, no CPU fallback.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, Dict, List, Tuple

import torch

from . import ops, synth


@dataclass
class FrameInputs:
    points: torch.Tensor      # [N,8] f32 (x,y,z,intensity,dt, no-aug xyz)
    mask: torch.Tensor        # [cams,classes,H,W] u8 instance-id planes
    lidar2img: torch.Tensor   # [cams,4,4] f32
    pt_feats: torch.Tensor    # [N,64] f32 stand-in for the VFE Linear outputs until the GEMM lands
    seg_logits: torch.Tensor  # [N,11]
    vote_preds: torch.Tensor  # [N,33]
    seg_feats: torch.Tensor   # [N,131]

    def nbytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in (self.points, self.mask, self.lidar2img))


def synth_frame_host(n: int, sweeps: int, seed: int, pin: bool = True) -> Dict[str, torch.Tensor]:
    """Host-side (pinned) buffers of one synthetic nuScenes-shaped frame."""
    import numpy as np

    pts = synth.ring_points(n, sweeps=sweeps, seed=seed)
    mask = synth.mask_planes(seed=seed)
    l2i = synth.lidar2img()
    out = dict(points=torch.from_numpy(pts), mask=torch.from_numpy(mask), lidar2img=torch.from_numpy(l2i))
    if pin and torch.cuda.is_available():
        out = {k: v.pin_memory() for k, v in out.items()}
    del np
    return out


def frame_to_device(host: Dict[str, torch.Tensor], dev: torch.device, seed: int = 0) -> FrameInputs:
    g = torch.Generator(device=dev).manual_seed(seed)
    n = host["points"].size(0)
    return FrameInputs(
        points=host["points"].to(dev, non_blocking=True),
        mask=host["mask"].to(dev, non_blocking=True),
        lidar2img=host["lidar2img"].to(dev, non_blocking=True),
        pt_feats=torch.randn(n, 64, device=dev, generator=g),
        seg_logits=torch.randn(n, 11, device=dev, generator=g),
        vote_preds=torch.randn(n, 33, device=dev, generator=g),
        seg_feats=torch.randn(n, 131, device=dev, generator=g),
    )


NUSC_GRID_ZYX = (40, 512, 512)
PRE_GRID_ZYX = (80, 1024, 1024)
PRE_VOXEL = (0.1, 0.1, 0.1)


def build_stages(inp: FrameInputs) -> Tuple[List[Tuple[str, Callable[[], None]]], Dict[str, torch.Tensor]]:
    """Returns ([(name, fn)], state).  Stage fns fill `state` in place."""
    st: Dict[str, torch.Tensor] = {}
    n = inp.points.size(0)

    def voxelize():
        st["coors"] = ops.voxelize(inp.points, synth.NUSC_VOXEL, synth.NUSC_RANGE, floor_mode=0)

    def rank():
        uniq, inv, _ = ops.unique_rows(st["coors"], lo=[0, 0, 0], ext=list(NUSC_GRID_ZYX), inv_dtype=torch.int32)
        st["voxel_coors"], st["inv"] = uniq, inv

    def csr():
        st["csr"] = ops.build_csr(st["inv"], st["voxel_coors"].size(0))

    def vfe_scatter():
        c = st["csr"]
        st["voxel_mean"] = ops.segment_reduce(inp.points[:, :5], c, "mean")
        st["vfe0"] = ops.segment_reduce(inp.pt_feats, c, "max")
        st["vfe1"] = ops.segment_reduce(inp.pt_feats, c, "max")

    def neck():
        out = torch.empty((n, 128), dtype=torch.float32, device=inp.points.device)
        ops.gather_rows(st["vfe0"], st["inv"], out=out[:, :64])
        ops.gather_rows(st["vfe1"], st["inv"], out=out[:, 64:])
        st["pt_voxel_feats"] = out

    def project():
        st["ids_sel"], st["cam_sel"], st["fg"] = ops.project_sample_select(inp.points[:, 5:8], inp.lidar2img, inp.mask)

    def pre_voxelize():
        coors = ops.voxelize(inp.points, PRE_VOXEL, synth.NUSC_RANGE, floor_mode=1)
        uniq, inv, _ = ops.unique_rows(coors, lo=[0, 0, 0], ext=list(PRE_GRID_ZYX), inv_dtype=torch.int32)
        c = ops.build_csr(inv, uniq.size(0))
        st["pre_coors"] = uniq
        st["pre_points"] = ops.segment_reduce(inp.points[:, :5], c, "mean")
        st["pre_logits"] = ops.segment_reduce(inp.seg_logits, c, "mean")
        st["pre_votes"] = ops.segment_reduce(inp.vote_preds, c, "mean")
        st["pre_feats"] = ops.segment_reduce(inp.seg_feats, c, "mean")
        st["pre_centers"] = ops.segment_reduce(inp.vote_preds, c, "mean")

    stages = [("voxelize", voxelize), ("rank", rank), ("csr", csr), ("vfe_scatter", vfe_scatter), ("neck", neck),
              ("project", project), ("pre_voxelize", pre_voxelize)]
    return stages, st


def algorithmic_bytes(inp: FrameInputs, st: Dict[str, torch.Tensor]) -> Dict[str, int]:
    """Algorithmic HBM bytes per stage (SURVEY.md §8d formulas; useful bytes once, no temporaries)."""
    n = inp.points.size(0)
    m = st["voxel_coors"].size(0)
    mp = st["pre_coors"].size(0)
    cams, classes = inp.mask.shape[:2]

    def scat(nn, c, mm):
        return 4 * nn * c + 4 * nn + 4 * mm * c

    return {
        "voxelize": 24 * n,
        "rank": 12 * n + 4 * n + 12 * m,
        "csr": 4 * n + 8 * n + 4 * (m + 1),
        "vfe_scatter": scat(n, 5, m) + 2 * scat(n, 64, m),
        "neck": 2 * (4 * n * 64 + 4 * n + 4 * m * 64),
        "project": (12 + cams * classes + 4 * classes + 2) * n,
        "pre_voxelize": 24 * n + 16 * n + 12 * mp + 12 * n + sum(scat(n, c, mp) for c in (5, 11, 33, 131, 33)),
    }
