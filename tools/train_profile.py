"""Device time per kernel of one training step of the segmentation stage (torch.profiler), 300 k points."""
import collections, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from torch.autograd import DeviceType
import bench
from fullysparsefusion_b200 import synth
from fullysparsefusion_b200.train import SegmentorTrainer
dev = torch.device("cuda:0")
pts = torch.from_numpy(synth.ring_points(300000, sweeps=10, seed=0)).to(dev)
lab, vote = bench.synth_labels(pts)
torch.manual_seed(0)
trainer = SegmentorTrainer(bench.make_model().to(dev))
for _ in range(3):
    trainer.step(pts, lab, vote)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    trainer.step(pts, lab, vote)
    torch.cuda.synchronize()
kern = collections.defaultdict(lambda: [0, 0.0])
for e in prof.events():
    if e.device_type == DeviceType.CUDA:
        kern[e.name][0] += 1
        kern[e.name][1] += e.device_time
tot = sum(v[1] for v in kern.values())
print("device busy ms", round(tot / 1e3, 2), "launches", sum(v[0] for v in kern.values()))
for k, v in sorted(kern.items(), key=lambda kv: -kv[1][1])[:25]:
    print(f"{v[1] / 1e3:8.2f} ms {v[0]:5d}  {k[:110]}")
