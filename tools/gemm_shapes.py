"""Per-shape table of every timed op of one FSF frame (CUDA events around each op, ops.DETAIL names).
Usage: python tools/gemm_shapes.py [points] [--scope full] > gpurun_out/gemm_shapes.txt"""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from fullysparsefusion_b200 import ops

points = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 300000
full = "--scope" in sys.argv and "full" in sys.argv
dev = torch.device("cuda:0")
f = {k: v.to(dev) for k, v in bench.synth_frame(points, 10, 0).items()}
model = bench.make_model().to(dev)
with torch.no_grad():
    st = model(f["points"], f["mask"], f["anno"], f["lidar2img"])
    bench.calibrate_seg_head(model, st["seg_logits"])
    def frame():
        stages, st = model.stages(f["points"], f["mask"], f["anno"], f["lidar2img"])
        for name, fn in stages:
            fn()
        if full:
            model.refine(st, f["points"])
            model.get_bboxes(st)
    for _ in range(3):
        frame()
    torch.cuda.synchronize()
    ops.DETAIL = True
    ops.PROFILER = []
    n = 5
    for _ in range(n):
        frame()
    torch.cuda.synchronize()
    prof, ops.PROFILER = ops.PROFILER, None
agg = collections.OrderedDict()
for name, a, b, nb, fl in prof:
    k = agg.setdefault(name, [0.0, 0, 0, 0])
    k[0] += a.elapsed_time(b); k[1] += 1; k[2] += bench.resolve(nb); k[3] += bench.resolve(fl)
tot = sum(v[0] for v in agg.values()) / n
print(f"timed ops: {tot:.3f} ms/frame")
for name, (ms, c, nb, fl) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{ms/n:8.3f} ms/frame {c//n:3d} calls {ms/c*1e3:8.1f} us/call {nb/(ms*1e-3)/1e9:8.1f} GB/s {fl/(ms*1e-3)/1e12:7.1f} TF/s  {name}")
