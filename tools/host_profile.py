"""Host-side (Python) cost of one FSF frame: cProfile over 5 frames, top functions by own time and by cumulative time."""
import os, sys, cProfile, pstats, io, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
dev = torch.device("cuda:0")
f = {k: v.to(dev) for k, v in bench.synth_frame(300000, 10, 0).items()}
model = bench.make_model().to(dev)
with torch.no_grad():
    st = model(f["points"], f["mask"], f["anno"], f["lidar2img"])
    bench.calibrate_seg_head(model, st["seg_logits"])
    for _ in range(3):
        model(f["points"], f["mask"], f["anno"], f["lidar2img"])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        model(f["points"], f["mask"], f["anno"], f["lidar2img"])
    torch.cuda.synchronize()
    print("frame ms (no profiler):", (time.perf_counter() - t0) / 5 * 1e3)
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(5):
        model(f["points"], f["mask"], f["anno"], f["lidar2img"])
    torch.cuda.synchronize()
    pr.disable()
for key in ("tottime", "cumtime"):
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats(key).print_stats(28)
    print("\n".join(l[:150] for l in s.getvalue().splitlines()[4:42]))
