import os, sys
sys.path.insert(0, "/root/repo")
import torch, torch.nn.functional as F
import bench
from fullysparsefusion_b200 import modules as M, ops, synth
dev = torch.device("cuda:0")
f = {k: v.to(dev) for k, v in bench.synth_frame(300000, 10, 0).items()}
pts = f["points"]
g = torch.Generator(device=dev).manual_seed(0)
c4 = F.pad(ops.voxelize(pts, synth.NUSC_VOXEL, synth.NUSC_RANGE, floor_mode=0), (1, 0), value=0)
plan = M.ScatterPlan(c4, lo=[0, 0, 0, 0], ext=[1, 40, 512, 512], want_index=True)
nbr = ops.conv_rulebook(plan.new_coors, plan.index, 3, 1, 1)
a = torch.randn(plan.m, 128, device=dev, generator=g)
w = ops.gemm_prepack(torch.randn(27, 128, 128, device=dev, generator=g) * 0.03)
order = ops.rulebook_row_order(nbr)
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
print("conv natural", t(lambda: ops.gather_gemm(a, w, nbr=nbr, act="relu")))
print("conv sorted ", t(lambda: ops.gather_gemm(a, w, nbr=nbr, act="relu", row_order=order)))
print("row_order   ", t(lambda: ops.rulebook_row_order(nbr)))
mask = ((nbr >= 0).long() << torch.arange(27, device=dev)[:, None]).sum(0)
def act(o):
    m = mask[o.long()]
    pad = (-len(m)) % 128
    m = torch.cat([m, m.new_zeros(pad)]).view(-1, 128)
    u = m[:, 0].clone()
    for j in range(1, 128): u |= m[:, j]
    return sum(((u >> k) & 1).sum().item() for k in range(27)) / len(u)
print("active offsets per tile: natural", act(torch.arange(len(mask), device=dev)), "sorted", act(order))
