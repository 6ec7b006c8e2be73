"""Kernel-level trace of one FSF frame (torch.profiler / CUPTI): GPU busy time vs elapsed per stage, top kernels.
Not a bench: profiler overhead inflates host time; use the busy/elapsed ratio and the kernel sums only."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity, record_function
import bench
dev = torch.device("cuda:0")
f = {k: v.to(dev) for k, v in bench.synth_frame(300000, 10, 0).items()}
model = bench.make_model().to(dev)
with torch.no_grad():
    st = model(f["points"], f["mask"], f["anno"], f["lidar2img"])
    bench.calibrate_seg_head(model, st["seg_logits"])
    for _ in range(3):
        st = model(f["points"], f["mask"], f["anno"], f["lidar2img"])
        model.refine(st, f["points"])
        model.get_bboxes(st)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(2):
            stages, st = model.stages(f["points"], f["mask"], f["anno"], f["lidar2img"])
            stages = stages + [("refine", lambda: model.refine(st, f["points"])), ("boxes", lambda: model.get_bboxes(st))]
            for name, fn in stages:
                with record_function("stage:" + name):
                    fn()
        torch.cuda.synchronize()
ev = prof.events()
kern = [e for e in ev if e.device_type == torch.autograd.DeviceType.CUDA]
kern.sort(key=lambda e: e.time_range.start)
stages_cpu = [e for e in ev if e.name.startswith("stage:") and e.device_type == torch.autograd.DeviceType.CPU]
agg = collections.defaultdict(lambda: [0.0, 0])
for e in kern:
    k = agg[e.name[:70]]
    k[0] += e.time_range.elapsed_us(); k[1] += 1
tot = sum(v[0] for v in agg.values())
span = kern[-1].time_range.end - kern[0].time_range.start
print(f"kernels+memcpy: {len(kern)} events, busy {tot/2e3:.2f} ms/frame, span {span/2e3:.2f} ms/frame")
for name, (us, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:45]:
    print(f"{us/2e3:8.3f} ms/frame {n//2:4d} launches  {name}")
# gaps: idle time between consecutive device events
gaps = []
for a, b in zip(kern[:-1], kern[1:]):
    g = b.time_range.start - a.time_range.end
    if g > 0:
        gaps.append((g, a.name[:40], b.name[:40]))
print("idle total %.2f ms/frame over %d gaps; gaps > 50us: %d" % (sum(g[0] for g in gaps) / 2e3, len(gaps), sum(1 for g in gaps if g[0] > 50)))
for g in sorted(gaps, reverse=True)[:25]:
    print("  gap %7.1f us after %-40s before %s" % g)
for s in stages_cpu:
    print(s.name, "cpu span %.2f ms" % (s.time_range.elapsed_us() / 1e3))
