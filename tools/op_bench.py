"""Op-level roofline of the scatter + projection family (the kernels BASELINE.json's metric names) at the frame size
(300 k points) and BASELINE configs[4] (1 M points): algorithmic bytes / CUDA-event time vs the measured
HBM peak.  Inputs rotate over buffers larger than L2.  `python tools/op_bench.py [points ...]` prints one JSON line;
bench.py imports run() for its `roofline_ops` table."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F


def _time(fn, reps=20, warm=3):
    for _ in range(warm):
        fn(0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3  # us


def run(points: int, peak_gbs: float, dev=None):
    from fullysparsefusion_b200 import modules as M, ops, synth
    dev = dev or torch.device("cuda:0")
    sweeps = max(1, round(points / 30000))
    g = torch.Generator(device=dev).manual_seed(0)
    pts = torch.from_numpy(synth.ring_points(points, sweeps=sweeps, seed=0)).to(dev)
    mask = torch.from_numpy(synth.mask_planes(seed=0)).to(dev)
    anno = torch.from_numpy(synth.mask_anno(mask.cpu().numpy(), seed=0)).to(dev)
    l2i = torch.from_numpy(synth.lidar2img()).to(dev)
    n = pts.size(0)
    out = {}

    def rec(name, us, nbytes):
        gbs = nbytes / us / 1e3
        out[name] = {"us": round(us, 1), "alg_MB": round(nbytes / 1e6, 1), "GB/s": round(gbs, 1), "frac": round(gbs / peak_gbs, 3)}

    rot = 3  # rotating copies: 3 x (n x 132 x 4 B) >= 475 MB at 300 k points > 126 MB L2
    # voxelize (24 B/pt)
    ptsr = [pts.clone() for _ in range(rot)]
    rec("voxelize", _time(lambda i: ops.voxelize(ptsr[i % rot], synth.NUSC_VOXEL, synth.NUSC_RANGE, floor_mode=0)), 24 * n)
    # scatter_v2 pieces on the 0.1 m pre-voxelization (the largest scatter traffic of the frame, §8 a13)
    c4 = F.pad(ops.voxelize(pts, (0.1, 0.1, 0.1), synth.NUSC_RANGE, floor_mode=1), (1, 0), value=0)
    plan = M.ScatterPlan(c4, lo=[0, 0, 0, 0], ext=[1, 80, 1024, 1024])
    m = plan.m
    for c in (132, 128, 64, 33):
        feats = []
        for _ in range(rot):
            f = ops.empty_rows(n, c, dev)
            f.copy_(torch.randn(n, c, device=dev, generator=g))
            feats.append(f)
        cw = feats[0].size(1)
        for mode in (("mean",) if c in (132, 33) else ("max",)):
            rec(f"segment_reduce[{mode},c={c}]", _time(lambda i: plan.reduce(feats[i % rot], mode)), 4 * n * cw + 8 * n + 4 * m * cw)
        if c == 128:
            vox = [torch.randn(m, c, device=dev, generator=g) for _ in range(rot)]
            dst = ops.empty_rows(n, c, dev)
            rec("gather_rows[c=128]", _time(lambda i: ops.gather_rows(vox[i % rot], plan.inv32, out=dst)), 4 * n * c + 8 * n + 4 * m * c)
        del feats
    # instance-level scatter_max (SIR: n rows over ~250 ids)
    ids = torch.randint(0, 248, (n, 1), device=dev, generator=g, dtype=torch.int32)
    plan_i = M.ScatterPlan(F.pad(ids, (2, 0), value=0))
    feats = [torch.randn(n, 128, device=dev, generator=g) for _ in range(rot)]
    rec("segment_reduce[max,c=128,ids=248]", _time(lambda i: plan_i.reduce(feats[i % rot], "max")), 4 * n * 128 + 8 * n + 4 * plan_i.m * 128)
    del feats
    # projection + sampling: drop-in contract ([N,6,10] i64: 552 B/pt) and the fused form the frame uses
    xyz = [pts[:, 5:8].contiguous().clone() for _ in range(rot)]
    rec("project_sample[drop-in i64]", _time(lambda i: ops.project_sample(xyz[i % rot], l2i, mask)), 552 * n)
    rec("project_sample_select[fused]", _time(lambda i: ops.project_sample_select(xyz[i % rot], l2i, mask, want_overlap=True, anno=anno,
                                                                                  anno_col=4, want_ids=False)), (12 + 60 + 2 + 1 + 40) * n)
    return {"points": n, "segments": m, "ops": out}


if __name__ == "__main__":
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
    except OSError:
        pass
    peak = peaks.get("hbm_gbs", 6650.0)
    sizes = [int(a) for a in sys.argv[1:]] or [300000, 1000000]
    print(json.dumps({"hbm_peak_gbs": peak, "runs": [run(p, peak) for p in sizes]}))
