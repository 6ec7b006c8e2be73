"""Gather-GEMM timing on the frame's real level-0 rulebook (160k voxels) under FSFB_GEMM_DEBUG switches.
  python tools/conv_real_experiments.py            # parent: loops over debug values in FSFB_DBG_LIST
"""
import os, sys, subprocess, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    import torch.nn.functional as F
    import bench
    from fullysparsefusion_b200 import modules as M, ops, synth
    dev = torch.device("cuda:0")
    f = {k: v.to(dev) for k, v in bench.synth_frame(300000, 10, 0).items()}
    pts = f["points"]
    g = torch.Generator(device=dev).manual_seed(0)
    c4 = F.pad(ops.voxelize(pts, synth.NUSC_VOXEL, synth.NUSC_RANGE, floor_mode=0), (1, 0), value=0)
    plan = M.ScatterPlan(c4, lo=[0, 0, 0, 0], ext=[1, 40, 512, 512], want_index=True)
    nbr = ops.conv_rulebook(plan.new_coors, plan.index, 3, 1, 1)
    order = ops.rulebook_row_order(nbr)
    res = {}
    m = plan.m
    def stages(order):
        o = order.long() if order is not None else torch.arange(m, device=dev)
        pad = (-m) % 128
        act = (nbr >= 0)[:, o]
        act = F.pad(act, (0, pad)).view(27, -1, 128)
        per_tile = act.any(-1).sum(0)          # active offsets per tile
        n_t = per_tile.numel()
        cta = torch.zeros(148, dtype=torch.long, device=dev)
        cta.index_add_(0, torch.arange(n_t, device=dev) % 148, per_tile)
        res.setdefault("per_cta", {})["sorted" if order is not None else "natural"] = dict(
            mean=float(cta.float().mean()), max=int(cta.max()), tile_max=int(per_tile.max()),
            tail=[int(x) for x in per_tile[-8:]], head=[int(x) for x in per_tile[:8]])
        return int(act.any(-1).sum()), int(act.sum())
    res["rows"] = m
    res["stages_sorted,pairs"] = stages(order)
    res["stages_natural"] = stages(None)[0]
    for cin, cout in ((128, 128), (64, 128)):
        a = torch.randn(m, cin, device=dev, generator=g)
        w = ops.gemm_prepack(torch.randn(27, cout, cin, device=dev, generator=g) * 0.03)
        for name, od in (("sorted", order), ("natural", None)):
            for _ in range(3):
                y = ops.gather_gemm(a, w, nbr=nbr, act="relu", row_order=od)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                y = ops.gather_gemm(a, w, nbr=nbr, act="relu", row_order=od)
            e1.record()
            torch.cuda.synchronize()
            res[f"{cin}x{cout}_{name}"] = round(e0.elapsed_time(e1) / 10, 4)
    print(json.dumps(res))
else:
    for dbg in [int(x) for x in os.environ.get("FSFB_DBG_LIST", "0").split(",")]:
        env = dict(os.environ, FSFB_GEMM_DEBUG=str(dbg))
        r = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True)
        print("debug", dbg, r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-800:], flush=True)
