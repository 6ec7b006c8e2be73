"""Generate tests/golden/*.npz by running the REFERENCE's own Python on seeded inputs.

Runs only in the build container (needs /root/reference).  The reference imports mmcv / mmdet /
mmdet3d / mmseg / torch_scatter / ingroup_indices / dynamic_point_pool_ext, none of which are
installable here, so those names are served by inert import stubs: registries whose
register_module() is the identity, base classes that are plain nn.Module, decorators that pass
through.  Three stubs carry behaviour, each a documented stand-in:
  * mmcv.cnn.build_norm_layer  → nn.LayerNorm / nn.BatchNorm1d (what mmcv builds for 'LN'/'BN1d')
  * torch_scatter.scatter_max / scatter → torch scatter_reduce_/index_add_ on CPU with the
    sequential first-max-wins argmax of torch-scatter's CPU kernel
  * mmdet3d.ops.Voxelization  → not called (the goldens use the in-tree torch.div formula)
Everything else that executes is the reference's code, unmodified, imported from
/root/reference: FSF.prj_points_2d / points_in_mask / frustum_gather / extract_fg_pts /
double_overlap_pts / get_cluster_delta_weighted / img_cross_attn's selection,
single_stage_fsd.find_connected_componets[_single_batch] (scipy), sst_ops.scatter_v2 /
build_mlp, Voxel2PointScatterNeck.forward, SSTInputLayer.get_inner_win_inds_slow,
VoteSegHead.decode_vote_targets.

Usage:  python tools/make_golden.py        (writes tests/golden/, prints a manifest)
"""
from __future__ import annotations

import importlib
import importlib.abc
import importlib.machinery
import os
import sys
import types

import numpy as np
import torch
from torch import nn

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
OUT = os.path.join(REPO, "tests", "golden")
sys.path.insert(0, REPO)

STUB_ROOTS = ("mmcv", "mmdet", "mmdet3d", "mmseg", "torch_scatter", "ingroup_indices",
              "dynamic_point_pool_ext", "torchex", "nuscenes", "shapely", "av2", "pyquaternion",
              "trimesh", "open3d", "numba", "lyft_dataset_sdk", "plyfile", "skimage", "ipdb", "terminaltables",
              "pycocotools", "tensorboard")


import abc


class _StubMeta(abc.ABCMeta):
    def __getattr__(cls, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _make_stub(name)

    def register_module(cls, *a, **k):  # REGISTRY.register_module() on the class itself
        return lambda c: c


class _Stub(nn.Module, metaclass=_StubMeta):
    """Subclassable, callable as a decorator factory, usable as a registry."""

    def __init__(self, *a, **k):
        super().__init__()

    def forward(self, *a, **k):
        if len(a) == 1 and (isinstance(a[0], type) or callable(a[0])):
            return a[0]  # decorator use
        raise RuntimeError("stub called")

    @classmethod
    def register_module(cls, *a, **k):  # REGISTRY.register_module() → identity decorator
        return lambda c: c

def _make_stub(name):
    return _StubMeta(name, (_Stub,), {})


class _StubModule(types.ModuleType):
    __path__: list = []

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        obj = _SPECIAL.get((self.__name__, name))
        if obj is None:
            obj = _make_stub(name)
        setattr(self, name, obj)
        return obj


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path, target=None):
        if fullname.split(".")[0] in STUB_ROOTS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        return _StubModule(spec.name)

    def exec_module(self, module):
        pass


# ---- behavioural stand-ins ---------------------------------------------------------------
def _build_norm_layer(cfg, num_features, postfix=""):
    t = cfg["type"]
    if t == "LN":
        return "ln", nn.LayerNorm(num_features, eps=cfg.get("eps", 1e-5))
    if t in ("BN1d", "naiveSyncBN1d", "BN"):
        return "bn", nn.BatchNorm1d(num_features, eps=cfg.get("eps", 1e-5), momentum=cfg.get("momentum", 0.1))
    raise NotImplementedError(t)


def _ts_scatter_max(src, index, dim=0, out=None, dim_size=None):
    assert dim == 0 and src.dim() == 2
    n, c = src.shape
    m = int(index.max()) + 1 if dim_size is None else dim_size
    res = torch.full((m, c), float("-inf"), dtype=src.dtype)
    arg = torch.full((m, c), n, dtype=torch.long)
    for i in range(n):  # torch-scatter CPU kernel: sequential, strict '>' keeps the first max
        s = int(index[i])
        upd = src[i] > res[s]
        res[s] = torch.where(upd, src[i], res[s])
        arg[s] = torch.where(upd, torch.full_like(arg[s], i), arg[s])
    res[arg == n] = 0
    return res, arg


def _ts_scatter(src, index, dim=0, out=None, dim_size=None, reduce="sum"):
    assert dim == 0 and src.dim() == 2
    m = int(index.max()) + 1 if dim_size is None else dim_size
    res = torch.zeros((m, src.size(1)), dtype=src.dtype).index_add_(0, index, src)
    if reduce in ("sum", "add"):
        return res
    if reduce == "mean":
        cnt = torch.zeros(m, dtype=src.dtype).index_add_(0, index, torch.ones_like(index, dtype=src.dtype))
        return res / cnt.clamp(min=1)[:, None]
    raise NotImplementedError(reduce)


def _multi_apply(func, *args, **kwargs):
    from functools import partial
    pfunc = partial(func, **kwargs) if kwargs else func
    return tuple(map(list, zip(*map(pfunc, *args))))


_SPECIAL = {
    ("mmcv.cnn", "build_norm_layer"): _build_norm_layer,
    ("torch_scatter", "scatter_max"): _ts_scatter_max,
    ("torch_scatter", "scatter"): _ts_scatter,
    ("mmdet.core", "multi_apply"): _multi_apply,
}


def import_reference():
    assert os.path.isdir(REF), "tools/make_golden.py needs /root/reference (build container only)"
    sys.meta_path.insert(0, _Finder())
    sys.path.insert(0, REF)
    sst_ops = importlib.import_module("projects.mmdet3d_plugin.ops.sst_ops")
    fsd = importlib.import_module("projects.mmdet3d_plugin.models.detectors.single_stage_fsd")
    fsf = importlib.import_module("projects.mmdet3d_plugin.models.detectors.FSF")
    neck = importlib.import_module("projects.mmdet3d_plugin.models.necks.voxel2point_neck")
    sst_in = importlib.import_module("projects.mmdet3d_plugin.models.middle_encoders.sst_input_layer")
    seg_head = importlib.import_module("projects.mmdet3d_plugin.models.decode_heads.segmentation_head")
    return dict(sst_ops=sst_ops, fsd=fsd, fsf=fsf, neck=neck, sst_in=sst_in, seg_head=seg_head)


def save(name, **arrays):
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **{k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v))
                                 for k, v in arrays.items()})
    print(f"{name}.npz  {os.path.getsize(path) / 1024:.1f} KiB  keys={sorted(arrays)}")


def main():
    from fullysparsefusion_b200 import synth  # shared seeded generators (also used by the GPU parity tests)

    ref = import_reference()
    torch.manual_seed(0)
    FSF = ref["fsf"].FSF
    fsf_self = types.SimpleNamespace()
    fsf_self.prj_points_2d = lambda *a: FSF.prj_points_2d(fsf_self, *a)
    fsf_self.points_in_mask = lambda *a: FSF.points_in_mask(fsf_self, *a)

    # ---- projection + nearest sampling (FSF.py:169-258) ------------------------------------
    for tag, (n, H, W, cams, classes, seed) in {
        "projection_small": (4000, 90, 160, 6, 10, 1),
        "projection_nusc": (3000, 900, 1600, 6, 10, 2),
    }.items():
        pts = synth.ring_points(n, sweeps=1, seed=seed)[:, :3]
        l2i = synth.lidar2img(cams, H, W)
        mask = synth.mask_planes(cams, classes, H, W, seed=seed)
        tp, tm, tl = torch.from_numpy(pts), torch.from_numpy(mask), torch.from_numpy(l2i)
        pts2d = FSF.prj_points_2d(fsf_self, tp, tl, H, W)
        ids = FSF.points_in_mask(fsf_self, tp, tm, tl)
        bidx = torch.zeros(n, dtype=torch.long)
        ids_fg = FSF.frustum_gather(fsf_self, bidx, tp, tm[None], None, [dict(lidar2img=l2i)])
        assert torch.equal(ids, ids_fg)
        # camera selection exactly as FSF.img_cross_attn (:714-718)
        cam_sel = ids.sum(-1).max(-1)[1]
        sel_mask = torch.nn.functional.one_hot(cam_sel, cams).bool().unsqueeze(-1)
        ids_sel = ids.masked_select(sel_mask).reshape(-1, classes)
        fg = ids.sum((-2, -1)) > 0
        extra = {} if tag == "projection_small" else {}
        save(tag, points=pts, lidar2img=l2i, mask_seed=seed, mask_shape=np.array(mask.shape),
             mask=mask if tag == "projection_small" else np.zeros(0, np.uint8),
             pts_2d=pts2d, ids=ids.to(torch.int16), cam_sel=cam_sel, ids_sel=ids_sel.to(torch.int16), fg=fg, **extra)

    # ---- frustum helpers: extract_fg_pts / double_overlap_pts / get_cluster_delta_weighted --
    n = 1500
    pts = synth.ring_points(n, sweeps=1, seed=5)[:, :3]
    l2i = synth.lidar2img(6, 90, 160)
    mask = synth.mask_planes(6, 10, 90, 160, seed=5, overlap=True)
    ids = FSF.points_in_mask(fsf_self, torch.from_numpy(pts), torch.from_numpy(mask), torch.from_numpy(l2i))
    g = torch.Generator().manual_seed(5)
    feat = torch.randn(n, 8, generator=g)
    w = torch.rand(n, generator=g)
    bz = torch.zeros(n, 1, dtype=torch.long)
    f1, b1, p1, o1, w1 = FSF.extract_fg_pts(fsf_self, feat, bz, torch.from_numpy(pts), ids, w)
    f2, b2, p2, o2, w2 = FSF.double_overlap_pts(fsf_self, f1, b1, p1, o1, w1)
    fsf_self.map_voxel_center_to_point = lambda a, b: FSF.map_voxel_center_to_point(fsf_self, a, b)
    sir_coors, _ = FSF.get_sir_coors(fsf_self, b2, o2, w2)
    delta, center, ccoors = FSF.get_cluster_delta_weighted(fsf_self, p2, sir_coors, w2.unsqueeze(-1))
    save("frustum_pool", points=pts, lidar2img=l2i, mask=mask, feat=feat, weights=w, ids=ids.to(torch.int16),
         fg_feat=f1, fg_points=p1, fg_ids=o1.to(torch.int16), fg_w=w1,
         dup_feat=f2, dup_points=p2, dup_obj=o2, dup_w=w2, sir_coors=sir_coors,
         f_cluster=delta, center=center, center_coors=ccoors)

    # ---- CCL (single_stage_fsd.py:45-82) -----------------------------------------------------
    fsd = ref["fsd"]
    for tag, (m, dist, seed, nb) in {"ccl_small": (300, 0.6, 3, 1), "ccl_mid": (2500, 0.5, 4, 1),
                                      "ccl_batch": (900, 0.7, 6, 3)}.items():
        cp, cb = synth.cluster_points(m, seed=seed, batches=nb)
        tp, tb = torch.from_numpy(cp), torch.from_numpy(cb)
        if nb == 1:
            lab = fsd.find_connected_componets_single_batch(tp, tb, dist)
            lab2 = fsd.find_connected_componets(tp, tb, dist)
            assert torch.equal(lab, lab2)
        else:
            lab = fsd.find_connected_componets(tp, tb, dist)
        save(tag, points=cp, batch_idx=cb, dist=np.float32(dist), labels=lab)

    # ---- scatter_v2 (sst_ops.py:150-177) ------------------------------------------------------
    sst_ops = ref["sst_ops"]
    g = torch.Generator().manual_seed(7)
    for tag, (n, c, spec) in {"scatter_c1": (10000, 128, ("ids", 32)), "scatter_vox": (6000, 16, ("vox", None)),
                              "scatter_odd": (3000, 5, ("vox", None)), "scatter_ties": (2000, 33, ("ids", 40))}.items():
        feat = torch.randn(n, c, generator=g)
        if tag == "scatter_c1":  # 5 MB of features: regenerate from a dedicated seed instead of storing
            feat = torch.randn(n, c, generator=torch.Generator().manual_seed(107))
        if tag == "scatter_ties":
            feat = torch.round(feat * 2) / 2  # many exact ties → argmax rule is exercised
        if spec[0] == "ids":
            coors = torch.stack([torch.zeros(n, dtype=torch.long), torch.zeros(n, dtype=torch.long),
                                 torch.randint(0, spec[1], (n,), generator=g)], 1)
        else:
            pts = torch.from_numpy(synth.ring_points(n, sweeps=1, seed=11)[:, :3])
            pc_range = torch.tensor([-51.2, -51.2, -5.0])
            vs = torch.tensor([0.2, 0.2, 0.2])
            cz = torch.div(pts - pc_range[None], vs[None], rounding_mode="floor").long()[:, [2, 1, 0]]
            coors = torch.cat([torch.zeros(n, 1, dtype=torch.long), cz], 1)
        out = {}
        for mode in ("max", "avg", "sum"):
            nf, nc, inv = sst_ops.scatter_v2(feat, coors, mode)
            out[f"out_{mode}"] = nf
        _, arg = _ts_scatter_max(feat, inv)
        import hashlib
        sha = np.frombuffer(hashlib.sha256(feat.numpy().tobytes()).digest(), dtype=np.uint8)
        save(tag, feat=feat if tag != "scatter_c1" else np.zeros(0, np.float32), feat_seed=107, feat_sha=sha,
             feat_shape=np.array(feat.shape), coors=coors.to(torch.int32), new_coors=nc.to(torch.int32),
             unq_inv=inv.to(torch.int32), argmax=arg.to(torch.int32), **out)

    # ---- voxel coordinates: the in-tree torch.div formula (single_stage_fsd.py:270,591-593) ----
    pts = synth.ring_points(20000, sweeps=2, seed=13)
    # add points that sit exactly on voxel boundaries (where floor(a/b) and div_floor can differ)
    edge = torch.arange(-250, 250, dtype=torch.float32)[:, None] * 0.2 + torch.zeros(1, 3)
    edge[:, 2] = (torch.arange(500, dtype=torch.float32) % 38) * 0.2 - 4.8
    allp = torch.cat([torch.from_numpy(pts[:, :3]), edge], 0)
    for tag, (rng, vs) in {"voxel_nusc": ([-51.2, -51.2, -5.0, 51.2, 51.2, 3.0], [0.2, 0.2, 0.2]),
                           "voxel_pre": ([-51.2, -51.2, -5.0, 51.2, 51.2, 3.0], [0.1, 0.1, 0.1])}.items():
        pc = torch.tensor(rng[:3])
        v = torch.tensor(vs)
        c1 = torch.div(allp - pc[None], v[None], rounding_mode="floor").long()[:, [2, 1, 0]]
        c0 = torch.floor((allp - pc[None]) / v[None]).long()[:, [2, 1, 0]]  # Voxelization kernel's rule
        save(tag, points=allp, pc_range=np.array(rng, np.float32), voxel_size=np.array(vs, np.float32),
             coors_divfloor=c1, coors_floor=c0)

    # ---- build_mlp (sst_ops.py:808-833) --------------------------------------------------------
    torch.manual_seed(3)
    for tag, (cin, dims, norm, act, is_head, rows) in {
        "mlp_ln_gelu": (48, [64, 64], dict(type="LN", eps=1e-3), "gelu", False, 257),
        "mlp_head": (64, [32, 32, 10], dict(type="LN", eps=1e-3), "gelu", True, 100),
        "mlp_bn_relu": (131, [128, 128], dict(type="naiveSyncBN1d", eps=1e-3, momentum=0.01), "relu", False, 300),
    }.items():
        mlp = sst_ops.build_mlp(cin, dims, norm, is_head=is_head, act=act)
        for mod in mlp.modules():  # non-trivial norm statistics / affine
            if isinstance(mod, nn.BatchNorm1d):
                mod.running_mean.normal_(0, 0.5)
                mod.running_var.uniform_(0.5, 2.0)
            if isinstance(mod, (nn.BatchNorm1d, nn.LayerNorm)):
                mod.weight.data.uniform_(0.5, 1.5)
                mod.bias.data.normal_(0, 0.2)
        mlp.eval()
        x = torch.randn(rows, cin)
        with torch.no_grad():
            y = mlp(x)
        sd = {k.replace(".", "__"): v for k, v in mlp.state_dict().items()}
        save(tag, x=x, y=y, **sd)

    # ---- Voxel2PointScatterNeck.forward (voxel2point_neck.py:27-70) --------------------------
    Neck = ref["neck"].Voxel2PointScatterNeck
    neck = Neck(point_cloud_range=[-51.2, -51.2, -5.0, 51.2, 51.2, 3.0], voxel_size=[0.2, 0.2, 0.2])
    neck.eval()
    pts = torch.from_numpy(synth.ring_points(2000, sweeps=1, seed=17)[:, :5])
    pc = torch.tensor([-51.2, -51.2, -5.0])
    cz = torch.div(pts[:, :3] - pc[None], torch.tensor([0.2, 0.2, 0.2])[None], rounding_mode="floor").long()[:, [2, 1, 0]]
    coors = torch.cat([torch.zeros(len(pts), 1, dtype=torch.long), cz], 1)
    vc, inv = torch.unique(coors, return_inverse=True, dim=0)
    vfeat = torch.randn(len(vc), 32)
    vfeat[::17] = -1  # padded (dropped) voxels, voxel_padding = -1
    res, pmask = neck(pts, coors, vfeat, inv, -1)
    save("neck", points=pts, coors=coors, voxel_feats=vfeat, voxel2point_inds=inv, out=res, mask=pmask)

    # ---- in-group indices slow oracle (sst_input_layer.py:200-208) ----------------------------
    L = ref["sst_in"].SSTInputLayer
    grp = torch.randint(0, 50, (4000,), generator=g)
    inner = L.get_inner_win_inds_slow(types.SimpleNamespace(), grp)
    save("ingroup", group=grp, inner=inner)

    # ---- vote decode (segmentation_head.py:265-266) ---------------------------------------------
    H = ref["seg_head"].VoteSegHead
    v = torch.randn(500, 33, generator=g)
    save("vote_decode", preds=v, offsets=H.decode_vote_targets(types.SimpleNamespace(), v))


def main_refine():
    """Goldens of the query-refinement stage's IN-TREE Python (written by `python tools/make_golden.py refine`):
      * BasePointBBoxCoder.decode + FSF.decode_stage_bboxes (core/bbox/coders/base_point_bbox_coder.py:59-82, FSF.py:1085-1095)
      * DynamicPointROIExtractor.forward with debug=True (roi_extractors/dynamic_point_roi_extractor.py:30-100) around the
        reference's DynamicPointPoolFunction (ops/dynamic_point_pool_op.py:7-57); the absent extension
        dynamic_point_pool_ext.forward is served by oracle.fsf_oracle.dynamic_point_pool — so the extractor's own
        assertions (:84-92) run on the oracle's semantics and the recorded outputs pin the glue (buffer protocol, valid-row
        filtering, field slicing)
      * FullySparseBboxHead.get_nonempty_roi_mask / align_roi_feature_and_rois (bbox_heads/fsd_bbox_head.py:152-197)."""
    from oracle import fsf_oracle as O

    import_reference()
    torch.manual_seed(0)
    coder_mod = importlib.import_module("projects.mmdet3d_plugin.core.bbox.coders.base_point_bbox_coder")
    fsf = importlib.import_module("projects.mmdet3d_plugin.models.detectors.FSF")
    pool_op = importlib.import_module("projects.mmdet3d_plugin.ops.dynamic_point_pool_op")
    ext_mod = importlib.import_module("projects.mmdet3d_plugin.models.roi_heads.roi_extractors.dynamic_point_roi_extractor")
    head_mod = importlib.import_module("projects.mmdet3d_plugin.models.roi_heads.bbox_heads.fsd_bbox_head")

    # ---- box decode ----
    g = torch.Generator().manual_seed(21)
    k = 400
    reg = torch.randn(k, 10, generator=g) * 0.7
    base = torch.randn(k, 3, generator=g) * 20
    coder = coder_mod.BasePointBBoxCoder(code_size=10)
    boxes = coder.decode(reg, base)
    self_ns = types.SimpleNamespace(bbox_coder=coder)
    rois = fsf.FSF.decode_stage_bboxes(self_ns, base, torch.zeros(k), [reg])
    save("box_decode", reg=reg, base=base, boxes=boxes, rois=rois)

    # ---- extractor glue around the pooling op ----
    def ext_forward(rois, pts, extra_wlh, max_inbox_point, out_pts_idx, out_roi_idx, out_pts_feats):
        p, r, f = O.dynamic_point_pool(rois.numpy(), pts.numpy(), extra_wlh, max_inbox_point, out_pts_idx.numel())
        out_pts_idx[: len(p)] = torch.from_numpy(p)
        out_roi_idx[: len(r)] = torch.from_numpy(r)
        out_pts_feats[: len(f)] = torch.from_numpy(f)

    pool_op.dynamic_point_pool_ext.forward = ext_forward
    rng = np.random.default_rng(22)
    n, kk = 4000, 30
    pts = np.concatenate([rng.uniform(-40, 40, (n, 2)), rng.uniform(-3, 1, (n, 1))], 1).astype(np.float32)
    centers = pts[rng.integers(0, n, kk)] + rng.normal(0, 0.3, (kk, 3)).astype(np.float32)
    dims = rng.uniform([1.5, 3.0, 1.2], [2.5, 6.0, 2.5], (kk, 3))
    rois7 = np.concatenate([centers, dims, rng.uniform(-np.pi, np.pi, (kk, 1))], 1).astype(np.float32)
    dense = (centers[:5, None, :] + rng.normal(0, 0.4, (5, 60, 3))).reshape(-1, 3).astype(np.float32)
    pts = np.concatenate([pts, dense])
    rois8 = np.concatenate([np.zeros((kk, 1), np.float32), rois7], 1)
    ext = ext_mod.DynamicPointROIExtractor(debug=True, extra_wlh=[1.0, 1.0, 1.0], max_inbox_point=512)
    inds, roi_inds, info = ext(torch.from_numpy(pts), torch.zeros(len(pts), dtype=torch.long), torch.from_numpy(rois8))
    save("roi_extractor", points=pts, rois=rois8, inds=inds, roi_inds=roi_inds, local_xyz=info["local_xyz"],
         boundary_offset=info["boundary_offset"], is_in_margin=info["is_in_margin"])
    # nothing pooled: the fake row
    far = rois8[:2].copy()
    far[:, 1:4] = 1e4
    ext_nd = ext_mod.DynamicPointROIExtractor(debug=False, extra_wlh=[1.0, 1.0, 1.0], max_inbox_point=512)   # the stock config's setting
    i2, r2, info2 = ext_nd(torch.from_numpy(pts), torch.zeros(len(pts), dtype=torch.long), torch.from_numpy(far))
    save("roi_extractor_empty", inds=i2, roi_inds=r2, local_xyz=info2["local_xyz"])

    # ---- RoI alignment ----
    Head = head_mod.FullySparseBboxHead
    hs = types.SimpleNamespace(training=False)
    feats = torch.randn(7, 12, generator=g)
    out_coors = torch.tensor([-1, 0, 2, 3, 7, 8, 11])
    mask = Head.get_nonempty_roi_mask(hs, out_coors, 13)
    aligned = Head.align_roi_feature_and_rois(hs, feats, out_coors, 13)
    save("roi_align", feats=feats, out_coors=out_coors, num_rois=13, mask=mask, aligned=aligned)


def main_wiring():
    """Goldens of the IN-TREE control flow around the un-vendored blocks (`python tools/make_golden.py wiring`): the reference's own
    SIR.forward (models/backbones/sir.py:65-85) and FullySparseBboxHead.forward (models/roi_heads/bbox_heads/fsd_bbox_head.py:
    95-151) run here, on CPU, with the registry types they build through builder.build_voxel_encoder ('SIRLayer', sir.py:41-62;
    'DynamicClusterVFE', fsd_bbox_head.py:62-87) served by the repo's OWN block classes — constructed from the reference's own
    kwargs, forward restated in plain torch below (the block's source is un-vendored; its parameters and state-dict names are the
    repo's).  What the goldens pin: which rows feed which block (points ‖ features ‖ f_cluster / 10), the order of the
    concatenated group features, use_middle_cluster_feature, the group coordinates handed to the RoI alignment, and that every
    kwarg the reference passes is accepted."""
    import torch.nn.functional as F

    from fullysparsefusion_b200 import modules as M
    from fullysparsefusion_b200.shims import registry_table

    class _CpuBlock:   # mixin: forward of modules.SIRLayer in stock torch ops (same arithmetic, CPU)
        def forward(self, features, coors, f_cluster=None, points=None, img_feats=None, img_metas=None, return_both=False,
                    unq_inv_once=None, new_coors_once=None):
            new_coors, inv = torch.unique(coors, return_inverse=True, dim=0)
            x = torch.cat([features[:, :3] / torch.tensor(self.xyz_normalizer), features[:, 3:]], 1)
            if self.with_rel_mlp:
                x = x * nn.Sequential.forward(self.rel_mlp, f_cluster / self.rel_dist_scaler)
            ori, cl, pf = x, [], None
            for i, vfe in enumerate(self.vfe_layers):
                y = vfe.norm(vfe.linear(x))
                pf = F.gelu(y) if vfe.act == "gelu" else F.relu(y)
                c = torch.full((new_coors.size(0), pf.size(1)), float("-inf")).scatter_reduce_(
                    0, inv[:, None].expand_as(pf), pf, reduce="amax", include_self=True)
                cl.append(c)
                if i != len(self.vfe_layers) - 1:
                    x = torch.cat([pf, c[inv]], 1)
            cluster = torch.cat(cl, 1)
            if return_both or self.return_point_feats:
                if self.with_shortcut and pf.shape == ori.shape:
                    pf = pf + ori
                return (pf, cluster, new_coors) if return_both else (pf, cluster)
            return cluster, new_coors

    table = {t: c for reg, t, c in registry_table() if reg == "VOXEL_ENCODERS"}
    cpu_types = {t: type("Cpu" + t, (_CpuBlock, c), {}) for t, c in table.items()}
    built = []

    def build_voxel_encoder(cfg):
        cfg = dict(cfg)
        cls = cpu_types[cfg.pop("type")]
        built.append((cls.__name__, sorted(cfg)))
        return cls(**cfg)

    _SPECIAL[("mmdet3d.models", "builder")] = types.SimpleNamespace(build_voxel_encoder=build_voxel_encoder)
    _SPECIAL[("mmdet3d.models.builder", "build_voxel_encoder")] = build_voxel_encoder   # when the submodule was imported first
    import_reference()
    sir_mod = importlib.import_module("projects.mmdet3d_plugin.models.backbones.sir")
    head_mod = importlib.import_module("projects.mmdet3d_plugin.models.roi_heads.bbox_heads.fsd_bbox_head")

    g = torch.Generator().manual_seed(31)
    torch.manual_seed(31)
    # ---- SIR backbone over (class, batch, cluster) ids ----
    n, cf = 1500, 13
    sir_kw = dict(num_blocks=3, in_channels=[3 + cf, 3 + 32, 3 + 32], feat_channels=[[32, 32]] * 3, rel_mlp_hidden_dims=[[16, 32]] * 3,
                  with_rel_mlp=True, norm_cfg=dict(type="LN", eps=1e-3), mode="max", xyz_normalizer=[20, 20, 4], act="gelu",
                  unique_once=True)
    sir = sir_mod.SIR(**sir_kw).eval()
    pts = torch.randn(n, 3, generator=g) * 10
    feats = torch.randn(n, cf, generator=g)
    coors = torch.stack([torch.randint(0, 3, (n,), generator=g), torch.zeros(n, dtype=torch.long), torch.randint(0, 40, (n,), generator=g)], 1)
    f_cluster = torch.randn(n, 3, generator=g)
    with torch.no_grad():
        out_feats, cluster_feats, out_coors = sir(pts, feats, coors, f_cluster)
    sd = {"sir_sd__" + k.replace(".", "__"): v for k, v in sir.state_dict().items()}
    save("wiring_sir", points=pts, features=feats, coors=coors, f_cluster=f_cluster, out_feats=out_feats, cluster_feats=cluster_feats,
         out_coors=out_coors, **sd)

    # ---- RoI head over pooled points (roi ids with the -1 fake group and empty rois) ----
    p, k, c0 = 900, 25, 16
    head_kw = dict(num_classes=10, num_blocks=3, in_channels=[3 + c0 + 13, 3 + 32 + 13, 3 + 32 + 13], feat_channels=[[32, 32]] * 3,
                   with_distance=False, with_cluster_center=False, with_rel_mlp=True, rel_mlp_hidden_dims=[[16, 32]] * 3,
                   rel_mlp_in_channels=[13] * 3, reg_mlp=None, cls_mlp=None, xyz_normalizer=[20, 20, 4], act="gelu", geo_input=True,
                   use_middle_cluster_feature=True, norm_cfg=dict(type="LN", eps=1e-3, momentum=0.01), unique_once=True)
    head = head_mod.FullySparseBboxHead(**head_kw).eval()
    xyz = torch.randn(p, 3, generator=g) * 10
    pf = torch.randn(p, c0, generator=g)
    roi_inds = torch.randint(0, k - 4, (p,), generator=g)       # the last rois stay empty
    roi_inds[:7] = -1                                             # rows of the extractor's fake group
    info = dict(local_xyz=torch.randn(p, 3, generator=g), boundary_offset=torch.rand(p, 6, generator=g),
                is_in_margin=(torch.rand(p, generator=g) < 0.3).float())
    rois = torch.cat([torch.zeros(k, 1), torch.randn(k, 3, generator=g) * 10, torch.rand(k, 3, generator=g) * 3 + 1,
                      torch.rand(k, 1, generator=g) * 6 - 3], 1)
    with torch.no_grad():
        roi_feats, nonempty = head(xyz, pf, info, roi_inds, rois)
    sd = {"head_sd__" + kk.replace(".", "__"): v for kk, v in head.state_dict().items()}
    save("wiring_roi_head", pts_xyz=xyz, pts_features=pf, roi_inds=roi_inds, rois=rois, local_xyz=info["local_xyz"],
         boundary_offset=info["boundary_offset"], is_in_margin=info["is_in_margin"], roi_feats=roi_feats, nonempty=nonempty, **sd)
    import json
    with open(os.path.join(OUT, "wiring_kwargs.json"), "w") as fh:   # the kwargs the reference passed to the registry builds
        json.dump(dict(sir=sir_kw, head=head_kw, built=built), fh, indent=1)
    print("registry builds:", [b[0] for b in built])


def main_groups():
    """Golden of the LiDAR-query clustering path (`python tools/make_golden.py groups`): the reference's own
    SingleStageFSD.group_sample (single_stage_fsd.py:802-865, with get_fg_mask :740-781, get_offset_weight :867-874,
    gather_group_by_names :891-904) followed by ClusterAssigner.forward (:921-982: filter_almost_empty, scatter_v2 'avg',
    scipy connected components, modify_cluster_by_class) for the six nuScenes class groups.  The in-group index of the
    'at least one point per sample' fallback comes from the reference's slow oracle get_inner_win_inds_slow."""
    ref = import_reference()
    fsd = ref["fsd"]
    S = fsd.SingleStageFSD
    L = ref["sst_in"].SSTInputLayer
    fsd.get_inner_win_inds = lambda x: L.get_inner_win_inds_slow(types.SimpleNamespace(), x)
    from fullysparsefusion_b200.fsf import NUSC
    cfg = dict(score_thresh=NUSC["score_thresh"], group_names=NUSC["group_names"], class_names=NUSC["class_names"], offset_weight="max")
    ns = types.SimpleNamespace(training=False, num_classes=10, test_cfg=cfg, train_cfg=None, cfg=cfg, runtime_info=None)
    for name in ("gather_group_by_names", "get_fg_mask", "get_sample_beg_position", "get_offset_weight"):
        setattr(ns, name, (lambda f: (lambda *a, **k: f(ns, *a, **k)))(getattr(S, name)))
    g = torch.Generator().manual_seed(31)
    n = 5000
    pts = torch.cat([torch.rand(n, 2, generator=g) * 80 - 40, torch.rand(n, 1, generator=g) * 4 - 3], 1)
    # clumps of foreground voxels so clusters exist; background logit dominant elsewhere
    logits = torch.randn(n, 11, generator=g)
    logits[:, 10] += 5.0
    for c, (lo, hi) in enumerate([(0, 400), (400, 600), (600, 700), (700, 1000), (1000, 1003), (1003, 1300)]):
        cls = NUSC["class_names"].index(NUSC["group_names"][c][0])
        logits[lo:hi, cls] += 9.0
        centre = pts[lo:hi].mean(0, keepdim=True)
        k = hi - lo
        pts[lo:hi] = centre + torch.randn(k, 3, generator=g) * torch.tensor([[1.5, 1.5, 0.3]])
    logits[:, NUSC["class_names"].index("barrier")] -= 12.0      # group 3 (barrier): no candidate at all → row-0 fallback
    pts[1000:1003] += torch.tensor([[0.0, 0.0, 0.0], [7.0, 0.0, 0.0], [0.0, 9.0, 0.0]])   # group 4: three isolated voxels → keep-all fallback
    offsets = torch.randn(n, 33, generator=g) * 0.2
    d = dict(seg_points=pts, seg_logits=logits, batch_idx=torch.zeros(n, dtype=torch.long))
    out = S.group_sample(ns, d, offsets)
    assigner = fsd.ClusterAssigner(cluster_voxel_size=[list(v) for v in NUSC["cluster_voxel_size"]], min_points=NUSC["min_points"],
                                   point_cloud_range=NUSC["point_cloud_range"], connected_dist=NUSC["connected_dist"],
                                   class_names=[f"g{i}" for i in range(6)], gpu_clustering=(False, False))
    assigner.num_classes = 6
    assigner.eval()
    cluster_inds_list, valid_mask_list = assigner(out["center_preds"], out["batch_idx"], origin_points=out["seg_points"])
    rows = [torch.nonzero(m).view(-1)[v] for m, v in zip(out["fg_mask_list"], valid_mask_list)]
    centers = [c[v] for c, v in zip(out["center_preds"], valid_mask_list)]
    save("group_cluster", points=pts, logits=logits, offsets=offsets, rows=torch.cat(rows), cluster_inds=torch.cat(cluster_inds_list),
         center_preds=torch.cat(centers), counts=np.array([len(r) for r in rows]),
         candidates=np.array([int(m.sum()) for m in out["fg_mask_list"]]))


def main_misc():
    """More of the hot path's IN-TREE Python, run as is (`python tools/make_golden.py misc`):
      * SingleStageFSD.pre_voxelize (single_stage_fsd.py:585-605)
      * FSF.get_single_cls_preds_2d + encode_preds_2d (FSF.py:449-504), as frustum_forward feeds encode_2d_mlp
      * FSF.img_cross_attn up to the MLP input: frustum_gather → camera select → get_all_cls_preds_2d → encode_preds_2d
        (FSF.py:694-728, 506-552)
      * FSF.get_point_fg_weights (FSF.py:345-355)
      * FSF.combine_frustum_and_fsd's index bookkeeping (FSF.py:657-692; the two MLPs are build_mlp, pinned separately)."""
    from fullysparsefusion_b200 import synth
    ref = import_reference()
    FSF, S, sst_ops = ref["fsf"].FSF, ref["fsd"].SingleStageFSD, ref["sst_ops"]
    g = torch.Generator().manual_seed(41)

    # ---- pre_voxelize ----
    n = 2000
    pts = torch.from_numpy(synth.ring_points(n, sweeps=3, seed=41)[:, :5])
    data = dict(seg_points=pts, seg_logits=torch.randn(n, 11, generator=g), seg_vote_preds=torch.randn(n, 33, generator=g),
                seg_feats=torch.randn(n, 19, generator=g), batch_idx=torch.zeros(n, dtype=torch.long))
    ns = types.SimpleNamespace(cfg=types.SimpleNamespace(pre_voxelization_size=(0.1, 0.1, 0.1)),
                               cluster_assigner=types.SimpleNamespace(point_cloud_range=[-51.2, -51.2, -5.0, 51.2, 51.2, 3.0]))
    vox = S.pre_voxelize(ns, data)
    save("pre_voxelize", points=pts, logits=data["seg_logits"], votes=data["seg_vote_preds"], feats=data["seg_feats"],
         v_points=vox["seg_points"], v_logits=vox["seg_logits"], v_votes=vox["seg_vote_preds"], v_feats=vox["seg_feats"])

    # ---- 2D prediction encodings ----
    H, W = 90, 160
    mask = synth.mask_planes(6, 10, H, W, seed=42, overlap=True)
    anno = synth.mask_anno(mask, seed=42)
    l2i = synth.lidar2img(6, H, W)
    fs = types.SimpleNamespace(num_classes=10, num_cams=6, encode_label_only=False, is_argo=False)
    for name in ("prj_points_2d", "points_in_mask", "frustum_gather", "get_all_cls_preds_2d", "encode_preds_2d", "encode_2d_feats",
                 "get_single_cls_preds_2d", "combine_by_batch"):
        setattr(fs, name, (lambda f: (lambda *a, **k: f(fs, *a, **k)))(getattr(FSF, name)))
    obj_coors = torch.stack([torch.zeros(40, dtype=torch.long), torch.zeros(40, dtype=torch.long),
                             torch.randint(0, int(anno[:, 7].max()) + 1, (40,), generator=g)], 1)
    preds = FSF.get_single_cls_preds_2d(fs, torch.from_numpy(anno)[None], obj_coors)
    enc = FSF.encode_preds_2d(fs, preds, W, H)
    save("encode_preds_2d", anno=anno, obj_coors=obj_coors, preds_2d=preds, enc=enc, img_w=W, img_h=H)

    npts = 3000
    p3 = torch.from_numpy(synth.ring_points(npts, sweeps=1, seed=43)[:, :3])
    captured = {}
    ident = lambda x: captured.setdefault("mlp_in", x.clone()) * 0 + x   # encode_mlp: record its input, pass it through
    out = FSF.img_cross_attn(fs, [p3], torch.zeros(npts, dtype=torch.long), torch.from_numpy(anno)[None], torch.from_numpy(mask)[None],
                             [dict(lidar2img=l2i)], ident)
    save("img_cross_attn", points=p3, lidar2img=l2i, mask=mask, anno=anno, scores=captured["mlp_in"])

    lg = torch.randn(500, 11, generator=g)
    save("fg_weights", logits=lg, weights=FSF.get_point_fg_weights(fs, lg))

    # ---- combine_frustum_and_fsd (identity MLPs: only the index bookkeeping is recorded) ----
    cs = types.SimpleNamespace(fsd_begin_idx=1000, combine_frustum_feat_mlp=lambda x: x, combine_fsd_feat_mlp=lambda x: x)
    kf, kl = 7, 11
    fr = dict(centers=torch.randn(kf, 3, generator=g), coors=torch.randint(0, 50, (kf, 3), generator=g), feats=torch.randn(kf, 5, generator=g),
              res=dict(cls_logits=[torch.randn(kf, 10, generator=g)], reg_preds=[torch.randn(kf, 10, generator=g)]),
              preds_2d=torch.randn(kf, 9, generator=g))
    ls = dict(centers=torch.randn(kl, 3, generator=g), coors=torch.randint(0, 50, (kl, 3), generator=g), feats=torch.randn(kl, 5, generator=g),
              res=dict(cls_logits=[torch.randn(kl, 10, generator=g)], reg_preds=[torch.randn(kl, 10, generator=g)]))
    oc, oco, ores, of, op2 = FSF.combine_frustum_and_fsd(cs, fr["centers"], fr["coors"], fr["res"], fr["feats"], fr["preds_2d"],
                                                         ls["centers"], ls["coors"], ls["res"], ls["feats"])
    save("combine", fr_centers=fr["centers"], fr_coors=fr["coors"], fr_cls=fr["res"]["cls_logits"][0], fr_reg=fr["res"]["reg_preds"][0],
         fsd_centers=ls["centers"], fsd_coors=ls["coors"], fsd_cls=ls["res"]["cls_logits"][0], fsd_reg=ls["res"]["reg_preds"][0],
         obj_centers=oc, obj_coors=oco, obj_cls=ores["cls_logits"][0], obj_reg=ores["reg_preds"][0], preds_2d=op2)


def main_heads():
    """The reference's own SparseClusterHeadV2 / FSDSeparateHead (dense_heads/sparse_cluster_head_v2.py:17-168,
    sparse_cluster_head.py:30-118) at reduced widths: state-dict keys (checkpoint compatibility of fsf.SparseClusterHeadV2) and
    forward outputs (`python tools/make_golden.py heads`).  mmdet's builders are stubs here, so build_head / build_bbox_coder are
    pointed at the in-tree classes they would resolve to."""
    import_reference()
    v2 = importlib.import_module("projects.mmdet3d_plugin.models.dense_heads.sparse_cluster_head_v2")
    sch = importlib.import_module("projects.mmdet3d_plugin.models.dense_heads.sparse_cluster_head")
    v2.builder.build_head = lambda cfg: v2.FSDSeparateHead(**{k: v for k, v in cfg.items() if k != "type"})
    sch.build_bbox_coder = lambda cfg: types.SimpleNamespace(code_size=cfg["code_size"])
    torch.manual_seed(51)
    names = ["car", "truck", "bus"]
    head = v2.SparseClusterHeadV2(
        num_classes=3, bbox_coder=dict(type="BasePointBBoxCoder", code_size=10), loss_cls=dict(type="FocalLoss"),
        loss_center=dict(type="L1Loss"), loss_size=dict(type="L1Loss"), loss_rot=dict(type="L1Loss"), in_channel=24,
        shared_mlp_dims=[32, 32], tasks=[dict(class_names=names)], class_names=names,
        common_attrs=dict(center=(3, 2, 16), dim=(3, 2, 16), rot=(2, 2, 16), vel=(2, 2, 16)), num_cls_layer=2, cls_hidden_dim=16,
        separate_head=dict(type="FSDSeparateHead", norm_cfg=dict(type="LN"), act="gelu"), norm_cfg=dict(type="LN"), act="relu")
    for mod in head.modules():
        if isinstance(mod, nn.LayerNorm):
            mod.weight.data.uniform_(0.5, 1.5)
            mod.bias.data.normal_(0, 0.2)
    head.eval()
    x = torch.randn(50, 24)
    with torch.no_grad():
        out = head(x)
    sd = {k.replace(".", "__"): v for k, v in head.state_dict().items()}
    save("cluster_head_v2", x=x, cls=out["cls_logits"][0], reg=out["reg_preds"][0], **sd)


def main_seghead():
    """The reference's own VoteSegHead (decode_heads/segmentation_head.py:16-104 over the in-tree Base3DDecodeHead,
    decode_heads/decode_head.py:9-106) at reduced widths, eval mode, non-trivial BN running statistics: state-dict keys (checkpoint
    compatibility of modules.VoteSegHead) and forward outputs (`python tools/make_golden.py seghead`)."""
    ref = import_reference()
    torch.manual_seed(71)
    head = ref["seg_head"].VoteSegHead(in_channel=24, num_classes=4, hidden_dims=[16, 16], dropout_ratio=0.0,
                                       norm_cfg=dict(type="naiveSyncBN1d"), act_cfg=dict(type="ReLU"),   # as FSF_nuScenes_config.py:78-95
                                       loss_decode=dict(type="CrossEntropyLoss", use_sigmoid=False, class_weight=[1.0] * 4 + [0.1],
                                                        loss_weight=10.0), loss_vote=dict(type="L1Loss", loss_weight=1.0))
    for mod in head.modules():
        if isinstance(mod, nn.BatchNorm1d):
            mod.running_mean.normal_(0, 0.3)
            mod.running_var.uniform_(0.5, 2.0)
            mod.weight.data.uniform_(0.5, 1.5)
            mod.bias.data.normal_(0, 0.2)
    head.eval()
    x = torch.randn(300, 24)
    with torch.no_grad():
        logits, votes = head(x)
    sd = {k.replace(".", "__"): v for k, v in head.state_dict().items()}
    save("vote_seg_head", x=x, logits=logits, votes=votes, num_classes=np.int64(head.num_classes), **sd)


def main_loading():
    """The reference's own LoadMaskFromFiles (datasets/pipelines/loading.py:22-339: load_nusc, load_argo + resize_img,
    load_waymo + resize_img_waymo, reorg_anno_*) on sample directories written in the format of
    tools/mask_tools/save_mask_nusc.py:138-171 (`python tools/make_golden.py loading`).  The sample directories are committed
    under tests/golden/mask_samples/ (they are the INPUT fixtures); the outputs go to tests/golden/mask_loading.npz — full
    planes for the small nuScenes sample, sha256 + strided rows for the full-size AV2 / Waymo stacks."""
    import copy
    import hashlib

    from fullysparsefusion_b200 import loading as L
    from fullysparsefusion_b200 import synth

    import_reference()
    ref = importlib.import_module("projects.mmdet3d_plugin.datasets.pipelines.loading")
    root = os.path.join(OUT, "mask_samples")
    out = {}

    # nuScenes layout: 6 cameras x 10 classes, no resize
    mask = synth.mask_planes(6, 10, 90, 160, seed=61, n_obj=120)
    anno = synth.mask_anno(mask, seed=61, n_obj=250)
    L.write_mask_sample(os.path.join(root, "nusc", "tok0"), mask, anno)
    L.write_mask_sample(os.path.join(root, "nusc", "empty"), np.zeros((6, 10, 45, 80), np.uint8), np.zeros((0, 9), np.float32))
    loader = ref.LoadMaskFromFiles(os.path.join(root, "nusc"))
    for tok in ("tok0", "empty"):
        res = loader(dict(sample_idx=tok))
        out[f"nusc_{tok}_mask"] = res["mask_data"]
        out[f"nusc_{tok}_anno"] = res["mask_anno"]
        assert res["mask_data"].dtype == torch.uint8

    def digest(t):
        return np.frombuffer(hashlib.sha256(t.contiguous().numpy().tobytes()).digest(), dtype=np.uint8)

    # AV2 layout: 7 single planes, the portrait front camera (2048 x 1550) resized to 1550 x 2048
    rng = np.random.default_rng(62)
    planes, rows = [], []
    oid = 1
    for cam in range(7):
        H, W = (2048, 1550) if cam == 0 else (1550, 2048)
        m = np.zeros((H, W), np.uint16)
        for _ in range(12):
            y0, x0 = int(rng.integers(0, H - 300)), int(rng.integers(0, W - 300))
            h, w = int(rng.integers(20, 300)), int(rng.integers(20, 300))
            m[y0:y0 + h, x0:x0 + w] = oid
            rows.append([x0 + 0.25, y0 + 0.5, x0 + w - 0.25, y0 + h - 0.5, float(rng.uniform(0.3, 1)), int(rng.integers(0, 26)), cam, oid, 1])
            oid += 1
        planes.append(m)
    sample = os.path.join(root, "argo", "uuid0")
    os.makedirs(sample, exist_ok=True)
    import cv2, json
    anno = [[] for _ in range(7)]
    for r in rows:
        anno[r[6]].append(dict(bbox=r[:4], score=r[4], category=r[5], cam_id=r[6], obj_id=r[7]))
    json.dump(anno, open(os.path.join(sample, "anno.json"), "w"), indent=2)
    for cam, m in enumerate(planes):
        cv2.imwrite(os.path.join(sample, f"{cam}.png"), m)
    l2i = [synth.lidar2img(7, 1550, 2048)[c].astype(np.float64) for c in range(7)]
    res = ref.LoadMaskFromFiles(os.path.join(root, "argo"), is_argo=True)(dict(img_info=dict(uuid="uuid0"), lidar2img=copy.deepcopy(l2i)))
    out.update(argo_sha=digest(res["mask_data"]), argo_shape=np.array(res["mask_data"].shape), argo_anno=res["mask_anno"],
               argo_l2i_in=np.stack(l2i), argo_l2i_out=np.stack(res["lidar2img"]), argo_rows=res["mask_data"][0, 0, ::97].clone(),
               argo_cols=res["mask_data"][0, 0, :, ::89].clone())
    assert res["mask_data"].dtype == torch.int32

    # Waymo layout: 5 cameras x 3 classes, side cameras 886 x 1920 resized to 1280 x 1920
    rows = []
    oid = 1
    sample = os.path.join(root, "waymo", "0001234")
    os.makedirs(sample, exist_ok=True)
    anno = [{n: [] for n in L.WAYMO_CLASSES} for _ in range(5)]
    for cam in range(5):
        H, W = (886, 1920) if cam >= 3 else (1280, 1920)
        for ci, name in enumerate(L.WAYMO_CLASSES):
            m = np.zeros((H, W), np.uint8)
            for _ in range(5):
                y0, x0 = int(rng.integers(0, H - 200)), int(rng.integers(0, W - 200))
                h, w = int(rng.integers(10, 200)), int(rng.integers(10, 200))
                m[y0:y0 + h, x0:x0 + w] = oid
                anno[cam][name].append(dict(bbox=[x0 + 0.5, y0 + 0.25, x0 + w - 0.5, y0 + h - 0.25], score=float(rng.uniform(0.3, 1)),
                                            category=ci, cam_id=cam, obj_id=oid))
                oid += 1
            cv2.imwrite(os.path.join(sample, f"{cam}_{name}.png"), m)
    json.dump(anno, open(os.path.join(sample, "anno.json"), "w"), indent=2)
    l2i = [synth.lidar2img(5, 1280, 1920)[c].astype(np.float64) for c in range(5)]
    res = ref.LoadMaskFromFiles(os.path.join(root, "waymo"), is_waymo=True)(
        dict(pts_filename="data/waymo/training/velodyne/0001234.bin", lidar2img=copy.deepcopy(l2i)))
    out.update(waymo_sha=digest(res["mask_data"]), waymo_shape=np.array(res["mask_data"].shape), waymo_anno=res["mask_anno"],
               waymo_l2i_in=np.stack(l2i), waymo_l2i_out=np.stack(res["lidar2img"]), waymo_rows=res["mask_data"][3:, :, ::61].clone())
    save("mask_loading", **out)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "heads":
        main_heads()
    elif len(sys.argv) > 1 and sys.argv[1] == "loading":
        main_loading()
    elif len(sys.argv) > 1 and sys.argv[1] == "seghead":
        main_seghead()
    elif len(sys.argv) > 1 and sys.argv[1] == "misc":
        main_misc()
    elif len(sys.argv) > 1 and sys.argv[1] == "wiring":
        main_wiring()
    elif len(sys.argv) > 1 and sys.argv[1] == "refine":
        main_refine()
    elif len(sys.argv) > 1 and sys.argv[1] == "groups":
        main_groups()
    else:
        main()
