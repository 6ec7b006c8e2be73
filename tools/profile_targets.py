"""Deterministic launches of the kernels judged in profiles/: run under ncu with -k <kernel> -s 2 -c 1.

  segreduce : pre_voxelize-shaped mean, 300k rows x 132 (131 + pad) channels over ~209k 0.1 m voxels
  project   : fused projection + sampling + camera select + score lookup, 300k points x 6 cams x 10 planes
  conv      : SubM 27-offset 128->128 gather-GEMM (A through TMEM) on the frame's 160k voxels
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
import bench
from fullysparsefusion_b200 import modules as M, ops, synth

dev = torch.device("cuda:0")
f = {k: v.to(dev) for k, v in bench.synth_frame(300000, 10, 0).items()}
pts = f["points"]
which = sys.argv[1] if len(sys.argv) > 1 else "all"
g = torch.Generator(device=dev).manual_seed(0)
if which in ("segreduce", "all"):
    c4 = F.pad(ops.voxelize(pts, (0.1, 0.1, 0.1), synth.NUSC_RANGE, floor_mode=1), (1, 0), value=0)
    plan = M.ScatterPlan(c4, lo=[0, 0, 0, 0], ext=[1, 80, 1024, 1024])
    feat = ops.empty_rows(pts.size(0), 131, dev)
    feat.copy_(torch.randn(pts.size(0), 131, device=dev, generator=g))
    for _ in range(3):
        out = plan.reduce(feat, "mean")
    torch.cuda.synchronize()
    print("segreduce rows", pts.size(0), "segments", plan.m, "alg bytes", 4 * pts.size(0) * 132 + 8 * pts.size(0) + 4 * plan.m * 132)
if which in ("project", "all"):
    for _ in range(3):
        res = ops.project_sample_select(pts[:, 5:8], f["lidar2img"], f["mask"], want_overlap=True, anno=f["anno"], want_ids=False)
    torch.cuda.synchronize()
    print("project points", pts.size(0), "alg bytes", pts.size(0) * (12 + 60 + 2 + 1 + 40))
if which in ("conv", "all"):
    c4 = F.pad(ops.voxelize(pts, synth.NUSC_VOXEL, synth.NUSC_RANGE, floor_mode=0), (1, 0), value=0)
    plan = M.ScatterPlan(c4, lo=[0, 0, 0, 0], ext=[1, 40, 512, 512], want_index=True)
    nbr = ops.conv_rulebook(plan.new_coors, plan.index, 3, 1, 1)
    a = torch.randn(plan.m, 128, device=dev, generator=g)
    w = ops.gemm_prepack(torch.randn(27, 128, 128, device=dev, generator=g) * 0.03)
    order = ops.rulebook_row_order(nbr) if os.environ.get("FSFB_NO_ROW_ORDER") != "1" else None
    for _ in range(3):
        y = ops.gather_gemm(a, w, nbr=nbr, act="relu", row_order=order)
    torch.cuda.synchronize()
    pairs = int((nbr >= 0).sum())
    print("conv voxels", plan.m, "pairs", pairs, "useful flops", 2 * pairs * 128 * 128)
