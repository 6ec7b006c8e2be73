"""Per-role cycle counters of the shared-memory-operand gather-GEMM (csrc/gemm_ss.cu) on the frame's level-0 rulebook and on
Linear shapes (FSFB_GEMM_TIMERS=1)."""
import os, sys, ctypes
os.environ["FSFB_GEMM_TIMERS"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.nn.functional as F
import bench
from fullysparsefusion_b200 import modules as M, ops, synth, _capi
dev = torch.device("cuda:0")
f = {k: v.to(dev) for k, v in bench.synth_frame(300000, 10, 0).items()}
pts = f["points"]
g = torch.Generator(device=dev).manual_seed(0)
c4 = F.pad(ops.voxelize(pts, synth.NUSC_VOXEL, synth.NUSC_RANGE, floor_mode=0), (1, 0), value=0)
plan = M.ScatterPlan(c4, lo=[0, 0, 0, 0], ext=[1, 40, 512, 512], want_index=True)
nbr = ops.conv_rulebook(plan.new_coors, plan.index, 3, 1, 1)
order = ops.rulebook_row_order(nbr)
nbr_ro = ops.permute_rulebook(nbr, order)
a = torch.randn(plan.m, 128, device=dev, generator=g)
a133 = ops.empty_rows(plan.m, 133, dev); a133.copy_(torch.randn(plan.m, 133, device=dev, generator=g))
w = ops.gemm_prepack(torch.randn(27, 128, 128, device=dev, generator=g) * 0.03)
lin_w = ops.gemm_prepack(torch.randn(1, 128, 128, device=dev, generator=g) * 0.03)
lin33 = ops.gemm_prepack(torch.randn(1, 33, 128, device=dev, generator=g) * 0.03)
lin133 = ops.gemm_prepack(torch.randn(1, 128, 133, device=dev, generator=g) * 0.03)
bias = torch.randn(128, device=dev, generator=g)
lib = _capi.load()
names = ["p_empty", "p_conv", "p_fetch", "p_stages", "mma_open", "mma_w", "mma_a", "mma_issue", "epi_pre", "epi_accfull", "epi_drain",
         "epi_out", "epi_units"]
print("rows", plan.m, "pairs/row", float((nbr >= 0).sum()) / plan.m)
for label, fn in (("split_rows only (timers below are stale)", lambda: ops.split_rows(a)),
                  ("conv sorted 128x128", lambda: ops.gather_gemm(a, w, nbr=nbr, act="relu", row_order=order, nbr_ro=nbr_ro)),
                  ("conv natural 128x128", lambda: ops.gather_gemm(a, w, nbr=nbr, act="relu")),
                  ("linear 128x128 bias relu", lambda: ops.gather_gemm(a, lin_w, bias=bias, act="relu")),
                  ("linear 128x128 plain", lambda: ops.gather_gemm(a, lin_w)),
                  ("linear 128x33", lambda: ops.gather_gemm(a, lin33)),
                  ("linear 133x128", lambda: ops.gather_gemm(a133, lin133)),
                  ("linear 128x128 ln gelu", lambda: ops.gather_gemm(a, lin_w, bias=bias, norm="ln", norm_w=bias, norm_b=bias, eps=1e-3, act="gelu"))):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    out = np.zeros((148, 32), dtype=np.uint32)
    rc = lib.fsfb_debug_gemm_ss_timers(out.ctypes.data_as(ctypes.c_void_p))
    if rc != 0:
        print(f"== {label}: {e0.elapsed_time(e1):.3f} ms")
        continue
    t = out.astype(np.float64)
    print(f"== {label}: {e0.elapsed_time(e1):.3f} ms; producer total clk mean {t[:,31].mean():.0f} max {t[:,31].max():.0f}")
    if t[:, 22].max() > 0:   # the row-tile Linear kernel (csrc/gemm_lin.cu): thread 0 of CTAs 1000..1147
        for i, n in enumerate(["prologue", "main loop", "last MMA", "phase 1", "statistics", "phase 2", "whole CTA"]):
            col = t[:, 16 + i]
            print(f"   lin {n:10s} mean {col.mean():10.0f}  min {col.min():10.0f}  max {col.max():10.0f}")
        continue
    for i, n in enumerate(names):
        col = t[:, i]
        print(f"   {n:12s} mean {col.mean():10.0f}  min {col.min():10.0f}  max {col.max():10.0f}")
