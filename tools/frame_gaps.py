"""Where a steady-state full-scope frame's wall time goes: warm kernel time vs GPU idle (host launch path + syncs).
torch.profiler (CUPTI) over 3 frames: per-frame wall (synchronised), sum of kernel durations, memcpy/memset time, the number
of launches and of blocking runtime calls.  Run under gpurun; prints one JSON object."""
import json, os, sys, time, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import bench
dev = torch.device("cuda:0")
f = {k: v.to(dev) for k, v in bench.synth_frame(300000, 10, 0).items()}
model = bench.make_model().to(dev)
with torch.no_grad():
    st = model(f["points"], f["mask"], f["anno"], f["lidar2img"])
    bench.calibrate_seg_head(model, st["seg_logits"])
    def frame():
        stages, st = model.stages(f["points"], f["mask"], f["anno"], f["lidar2img"])
        for _, fn in stages:
            fn()
        model.refine(st, f["points"])
        model.get_bboxes(st)
    for _ in range(3):
        frame()
    torch.cuda.synchronize()
    walls = []
    for _ in range(5):
        t0 = time.perf_counter(); frame(); torch.cuda.synchronize(); walls.append((time.perf_counter() - t0) * 1e3)
    n = 3
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(n):
            frame()
        torch.cuda.synchronize()
ev = prof.events()
kern = collections.defaultdict(lambda: [0, 0.0])
rt = collections.defaultdict(lambda: [0, 0.0])
from torch.autograd import DeviceType
for e in ev:
    if e.device_type == DeviceType.CUDA:
        kern[e.name][0] += 1; kern[e.name][1] += e.device_time if hasattr(e, "device_time") else e.cuda_time
    elif e.name.startswith("cuda") or e.name.startswith("cu"):
        rt[e.name][0] += 1; rt[e.name][1] += e.cpu_time
ktot = sum(v[1] for v in kern.values()) / n / 1e3
out = {"wall_ms_per_frame": [round(w, 2) for w in walls], "device_busy_ms_per_frame": round(ktot, 2),
       "device_launches_per_frame": sum(v[0] for v in kern.values()) / n,
       "runtime_calls_per_frame": {k: [v[0] / n, round(v[1] / n / 1e3, 2)] for k, v in sorted(rt.items(), key=lambda kv: -kv[1][1])[:12]},
       "top_device": {k[:70]: [v[0] / n, round(v[1] / n / 1e3, 3)] for k, v in sorted(kern.items(), key=lambda kv: -kv[1][1])[:30]}}
print(json.dumps(out, indent=1))
