"""The stock-PyTorch-on-CUDA "library bar" (BASELINE.md B2): the compositions the reference's hot functions
reduce to when its un-vendored extensions are replaced by what torch 2.x ships, timed on the SAME B200 next to the repo's kernels.

  scatter_v2      torch.unique(dim=0, return_inverse) + scatter_reduce_('amax') / index_add_ (ops/sst_ops.py:150-177)
  projection      mask.float() + [x,y,z,1] @ P^T + 6 x F.grid_sample(mode='nearest')      (models/detectors/FSF.py:169-226)
  gather          voxel_feats[voxel2point_inds]                                         (necks/voxel2point_neck.py:42-67)
  CCL             dense dist matrix on the GPU, D2H, scipy connected_components, H2D       (single_stage_fsd.py:45-82)
  SubM conv       per offset: index_select -> matmul (TF32 off, fp32) -> index_add_       (the classic spconv gather-GEMM-scatter)
  Linear          torch.nn.functional.linear + layer_norm + gelu                           (ops/sst_ops.py:808-833)

`python tools/library_bar.py [points]` prints one JSON object; bench.py imports run() for its `library_baseline` entry.  Nothing
here is on the product path."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.nn.functional as F


def _time(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3  # us


def run(points: int = 300000, dev=None):
    from scipy.sparse import csr_matrix
    from scipy.sparse.csgraph import connected_components

    from fullysparsefusion_b200 import modules as M, ops, synth
    dev = dev or torch.device("cuda:0")
    torch.backends.cuda.matmul.allow_tf32 = False       # the reference's fp32 arithmetic
    torch.backends.cudnn.allow_tf32 = False
    sweeps = max(1, round(points / 30000))
    g = torch.Generator(device=dev).manual_seed(0)
    pts = torch.from_numpy(synth.ring_points(points, sweeps=sweeps, seed=0)).to(dev)
    mask = torch.from_numpy(synth.mask_planes(seed=0)).to(dev)
    l2i = torch.from_numpy(synth.lidar2img()).to(dev)
    n = pts.size(0)
    out = {}

    def rec(name, lib_us, own_us):
        out[name] = {"library_us": round(lib_us, 1), "b200_us": round(own_us, 1), "speedup": round(lib_us / max(own_us, 1e-3), 2)}

    # ---- scatter_v2 on the 0.2 m voxels: unique + max / mean over 128 and 5 channels ----
    c3 = ops.voxelize(pts, synth.NUSC_VOXEL, synth.NUSC_RANGE, floor_mode=0)
    c4 = F.pad(c3, (1, 0), value=0)
    c4l = c4.long()
    feat = torch.randn(n, 128, device=dev, generator=g)

    def lib_scatter(mode, x):
        new_coors, inv = torch.unique(c4l, return_inverse=True, dim=0)
        m = new_coors.size(0)
        if mode == "max":
            o = torch.full((m, x.size(1)), float("-inf"), device=dev)
            return o.scatter_reduce_(0, inv[:, None].expand_as(x), x, reduce="amax", include_self=True)
        o = torch.zeros((m, x.size(1)), device=dev).index_add_(0, inv, x)
        return o / torch.bincount(inv, minlength=m).clamp(min=1)[:, None]

    def own_scatter(mode, x):
        plan = M.ScatterPlan(c4, lo=[0, 0, 0, 0], ext=[1, 40, 512, 512])
        return plan.reduce(x, mode)

    rec("scatter_v2[max,c=128] incl. unique", _time(lambda: lib_scatter("max", feat)), _time(lambda: own_scatter("max", feat)))
    rec("scatter_v2[mean,c=5] incl. unique", _time(lambda: lib_scatter("mean", pts[:, :5])), _time(lambda: own_scatter("mean", pts[:, :5].contiguous())))
    # ---- voxel -> point gather ----
    plan = M.ScatterPlan(c4, lo=[0, 0, 0, 0], ext=[1, 40, 512, 512])
    vox = torch.randn(plan.m, 128, device=dev, generator=g)
    inv64 = plan.inv32.long()
    rec("gather voxel->point [c=128]", _time(lambda: vox[inv64]), _time(lambda: ops.gather_rows(vox, plan.inv32)))
    # ---- projection + nearest sampling, the reference's literal op order (FSF.py:169-226) ----
    xyz = pts[:, 5:8].contiguous()

    def lib_project():
        mask_tensor = mask.float()
        cams, classes, h, w = mask_tensor.shape
        p4 = torch.cat([xyz, torch.ones_like(xyz[:, :1])], 1)
        pts_2d = torch.einsum("nk,cjk->cnj", p4, l2i)
        depth = pts_2d[..., 2:3]
        valid = depth > 1e-3
        uv = pts_2d[..., :2] / depth.clamp(1e-5, 1e5)
        uv = torch.stack([uv[..., 0] / w, uv[..., 1] / h], -1)
        uv = (uv - 0.5) * 2
        valid = valid & (uv[..., 0:1] > -1) & (uv[..., 0:1] < 1) & (uv[..., 1:2] > -1) & (uv[..., 1:2] < 1)
        uv = torch.where(valid, uv, torch.full_like(uv, -2.0))
        ids = [F.grid_sample(mask_tensor[c:c + 1], uv[c][None, None], mode="nearest", align_corners=False).squeeze(2).long()
               for c in range(cams)]
        return torch.cat(ids, 0).permute(2, 0, 1)

    rec("points_in_mask [N,6,10] i64", _time(lib_project, reps=3, warm=1), _time(lambda: ops.project_sample(xyz, l2i, mask)))
    # ---- CCL of one class group: dense matrix + scipy, as the live reference path (single_stage_fsd.py:69-82) ----
    ctr, _ = synth.cluster_points(4000, seed=3)
    tc = torch.from_numpy(ctr).to(dev)

    def lib_ccl():
        d = ((tc[:, None, :2] - tc[None, :, :2]) ** 2).sum(2) ** 0.5
        adj = (d < 0.6).cpu().numpy()
        _, lab = connected_components(csr_matrix(adj), directed=False)
        return torch.from_numpy(lab).to(dev)

    rec("connected components [4000 centres]", _time(lib_ccl, reps=3, warm=1), _time(lambda: ops.connected_components(tc, None, 0.6)))
    # ---- SubM 3x3x3 convolution 128 -> 128 on the frame's voxels: gather - GEMM - scatter per offset ----
    index = M.ScatterPlan(c4, lo=[0, 0, 0, 0], ext=[1, 40, 512, 512], want_index=True)
    nbr = ops.conv_rulebook(index.new_coors, index.index, 3, 1, 1)
    order = ops.rulebook_row_order(nbr)
    nbr_ro = ops.permute_rulebook(nbr, order)   # built once per rulebook, as modules.Rulebook does (not timed, like the rulebook)
    m = index.m
    a = torch.randn(m, 128, device=dev, generator=g)
    w = torch.randn(27, 128, 128, device=dev, generator=g) * 0.03
    pw = ops.gemm_prepack(w)
    pairs = [(torch.nonzero(nbr[k] >= 0)[:, 0], nbr[k][nbr[k] >= 0].long()) for k in range(27)]   # rulebook build is not timed

    def lib_conv():
        o = torch.zeros(m, 128, device=dev)
        for k in range(27):
            dst, src = pairs[k]
            if dst.numel():
                o.index_add_(0, dst, a.index_select(0, src) @ w[k].t())
        return torch.relu(o)

    rec(f"SubMConv3d 27x128->128 [{m} voxels]", _time(lib_conv, reps=3, warm=1), _time(lambda: ops.gather_gemm(a, pw, nbr=nbr, act="relu", row_order=order, nbr_ro=nbr_ro)))
    # ---- Linear -> LayerNorm -> GELU over the points ----
    x = torch.randn(n, 128, device=dev, generator=g)
    lw = torch.randn(128, 128, device=dev, generator=g) * 0.05
    nw, nb = torch.ones(128, device=dev), torch.zeros(128, device=dev)
    plw = ops.gemm_prepack(lw)
    rec("Linear+LN+GELU 128->128 [points]", _time(lambda: F.gelu(F.layer_norm(F.linear(x, lw), (128,), nw, nb, 1e-3))),
        _time(lambda: ops.gather_gemm(x, plw, norm="ln", norm_w=nw, norm_b=nb, eps=1e-3, act="gelu")))
    lib_total = sum(v["library_us"] for v in out.values())
    own_total = sum(v["b200_us"] for v in out.values())
    return {"points": n, "voxels": int(m), "ops": out, "sum_library_us": round(lib_total, 1), "sum_b200_us": round(own_total, 1),
            "note": "torch %s CUDA ops on the same GPU (TF32 off), one call of each op class; not a frame" % torch.__version__}


if __name__ == "__main__":
    print(json.dumps(run(int(sys.argv[1]) if len(sys.argv) > 1 else 300000), indent=1))
