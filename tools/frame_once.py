"""One full-scope frame inside an NVTX range for `ncu --nvtx --nvtx-include "frame/"` (per-launch device times of a steady-state
frame: the launch list committed under profiles/)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
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
    for _ in range(2):
        frame()
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_push("frame")
    frame()
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_pop()
print("frame done")
