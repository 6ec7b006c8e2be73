"""Deterministic launches of the gather-GEMM for ncu: level-0 SubM 27x128->128 on the synthetic frame's voxels (row-ordered)
and a Linear 128->128 over the same rows.  ncu ... -k regex:k_gather_gemm python tools/ss_profile_target.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
import bench
from fullysparsefusion_b200 import modules as M, ops, synth
dev = torch.device("cuda:0")
f = {k: v.to(dev) for k, v in bench.synth_frame(300000, 10, 0).items()}
g = torch.Generator(device=dev).manual_seed(0)
c4 = F.pad(ops.voxelize(f["points"], synth.NUSC_VOXEL, synth.NUSC_RANGE, floor_mode=0), (1, 0), value=0)
plan = M.ScatterPlan(c4, lo=[0, 0, 0, 0], ext=[1, 40, 512, 512], want_index=True)
nbr = ops.conv_rulebook(plan.new_coors, plan.index, 3, 1, 1)
order = ops.rulebook_row_order(nbr)
nbr_ro = ops.permute_rulebook(nbr, order)
a = torch.randn(plan.m, 128, device=dev, generator=g)
w = ops.gemm_prepack(torch.randn(27, 128, 128, device=dev, generator=g) * 0.03)
lin_w = ops.gemm_prepack(torch.randn(1, 128, 128, device=dev, generator=g) * 0.03)
bias = torch.randn(128, device=dev, generator=g)
for _ in range(2):
    ops.gather_gemm(a, w, nbr=nbr, act="relu", row_order=order, nbr_ro=nbr_ro)
    ops.gather_gemm(a, lin_w, bias=bias, act="relu")
torch.cuda.synchronize()
print("done", plan.m)
