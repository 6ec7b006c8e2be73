"""Fused projection: planar [cams,classes,H,W] planes against class-interleaved [cams,H,W,16] planes (experimental kernel), CUDA
events, 300 k and 1 M points at nuScenes size.  Algorithmic bytes per point as in bench.py's `ops` table (12 + 60 + outputs).

    python tools/op_bench_hwc.py      (GPU box)"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fullysparsefusion_b200 import ops, synth  # noqa: E402

dev = torch.device("cuda:0")
H, W = 900, 1600
mask = synth.mask_planes(6, 10, H, W, seed=0)
planar = torch.from_numpy(mask).to(dev)
hwc = torch.zeros((6, H, W, 16), dtype=torch.uint8, device=dev)
hwc[..., :10] = planar.permute(0, 2, 3, 1)
anno = torch.from_numpy(synth.mask_anno(mask, seed=0)).to(dev)
l2i = torch.from_numpy(synth.lidar2img(6, H, W)).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
out = {}
for n in (300_000, 1_000_000):
    xyz = torch.from_numpy(synth.ring_points(n, sweeps=10, seed=1)[:, :3].copy()).to(dev)
    want = ops.project_sample_select(xyz, l2i, planar, want_overlap=True, anno=anno)
    got = ops.project_sample_select_hwc(xyz, l2i, hwc, 10, want_overlap=True, anno=anno)
    same = all(torch.equal(a, b) for a, b in zip(want, got))
    bytes_pt = 12 + 60 + 40 + 3 + 40
    for name, fn in (("planar", lambda: ops.project_sample_select(xyz, l2i, planar, want_overlap=True, anno=anno)),
                     ("hwc16", lambda: ops.project_sample_select_hwc(xyz, l2i, hwc, 10, want_overlap=True, anno=anno))):
        ts = []
        for _ in range(20):
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ms = float(np.median(ts[5:]))
        out[f"{name}_{n}"] = {"ms": ms, "GB/s": n * bytes_pt / ms * 1e-6, "identical": same}
print(json.dumps(out, indent=1))
