"""Per-shape time of the sparse convolutions in one FSF frame (B200)."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from fullysparsefusion_b200 import ops
dev = torch.device("cuda:0")
f = {k: v.to(dev) for k, v in bench.synth_frame(300000, 10, 0).items()}
model = bench.make_model().to(dev)
with torch.no_grad():
    st = model(f["points"], f["mask"], f["anno"], f["lidar2img"])
    bench.calibrate_seg_head(model, st["seg_logits"])
    for _ in range(2):
        model(f["points"], f["mask"], f["anno"], f["lidar2img"])
    torch.cuda.synchronize()
    ops.PROFILER, ops.DETAIL = [], True
    n = 5
    for _ in range(n):
        model(f["points"], f["mask"], f["anno"], f["lidar2img"])
    torch.cuda.synchronize()
agg = collections.OrderedDict()
for name, a, b, nb, fl in ops.PROFILER:
    if "gemm" not in name:
        continue
    k = agg.setdefault(name, [0.0, 0, 0])
    k[0] += a.elapsed_time(b); k[1] += 1; k[2] += bench.resolve(fl)
for name, (ms, calls, fl) in agg.items():
    print(f"{name:32s} calls/frame {calls // n:3d}  ms/frame {ms / n:8.3f}  ms/call {ms / calls:7.3f}  useful TFLOP/s {fl / (ms * 1e-3) / 1e12:7.1f}")
