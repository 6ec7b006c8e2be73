"""End-to-end parity of one FSF frame (fullysparsefusion_b200.fsf.FSF) against the CPU oracle, stage by stage.

Each stage's oracle is fed the GPU path's inputs to that stage ("teacher forcing"), so a threshold flip in
one stage cannot cascade: discrete outputs (voxel/cluster indices, foreground rows, object ids, CCL
labels) are compared bit-exact, features within 1e-4 relative."""
import numpy as np
import pytest
import torch

from fullysparsefusion_b200 import fsf as FSFM
from fullysparsefusion_b200 import synth
from oracle import fsf_oracle as O
from oracle import fsf_oracle_frame as OF
from oracle import fsf_oracle_models as OM
from tests.test_gpu_modules import randomize, sd_np

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-4, 5e-5


def N(t):
    return t.detach().cpu().numpy()


def sub(sd, prefix):
    return {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}


def close(a, b, rtol=RTOL, atol=ATOL):
    np.testing.assert_allclose(N(a) if torch.is_tensor(a) else a, b, rtol=rtol, atol=atol)


def head_oracle(x, sd, prefix):
    """SparseClusterHeadV2.forward: shared MLP (LN, eps 1e-5, relu) + 5 task MLPs (LN, gelu, head)."""
    h = O.mlp_from_state_dict(x, sub(sd, prefix + "shared_mlp."), "ln", "relu", 1e-5)
    out = {k: O.mlp_from_state_dict(h, sub(sd, f"{prefix}task_heads.0.{k}."), "ln", "gelu", 1e-5)
           for k in ("center", "dim", "rot", "vel", "score")}
    return out["score"], np.concatenate([out["center"], out["dim"], out["rot"], out["vel"]], 1)


@pytest.fixture(scope="module")
def frame(cuda):
    n, H, W = 2500, 90, 160
    pts = synth.ring_points(n, sweeps=2, seed=21)
    mask = synth.mask_planes(6, 10, H, W, seed=21, overlap=True)
    anno = synth.mask_anno(mask, seed=21)
    l2i = synth.lidar2img(6, H, W)
    torch.manual_seed(0)
    model = randomize(FSFM.FSF(), seed=1)
    with torch.no_grad():  # un-zero the zero-initialised enhancement head so the stage is exercised
        model.segmentor_updated_mlp[-1].weight.normal_(0, 0.05)
        model.segmentor_updated_mlp[-1].bias.normal_(0, 0.05)
        model.segmentation_head.conv_seg.bias.copy_(torch.linspace(-1.0, 1.0, 11))
    model = model.to(cuda)
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda)
    st = model(T(pts), T(mask), T(anno), T(l2i))
    torch.cuda.synchronize()
    return dict(pts=pts, mask=mask, anno=anno, l2i=l2i, st=st, sd=sd_np(model), model=model)


def test_segment_stage(frame):
    st, sd, pts = frame["st"], frame["sd"], frame["pts"]
    cfg = FSFM.NUSC
    coors = np.concatenate([np.zeros((len(pts), 1), np.int64),
                            O.voxelize(pts, cfg["seg_voxel_size"], cfg["point_cloud_range"], 0).astype(np.int64)], 1)
    assert np.array_equal(N(st["coors4"]), coors)
    vf, vc, inv = OM.dynamic_scatter_vfe(pts[:, :5], coors, sub(sd, "voxel_encoder."), cfg["seg_voxel_size"], cfg["point_cloud_range"])
    assert np.array_equal(N(st["voxel_coors"]), vc) and np.array_equal(N(st["voxel2point_inds"]), inv)
    close(st["vfe_feats"], vf)
    m = frame["model"].backbone_unet
    want, _, _ = OM.simple_sparse_unet(N(st["vfe_feats"]), vc, sub(sd, "backbone_unet."), cfg["sparse_shape"], m.encoder_channels,
                                       m.encoder_paddings, ((512, 512, 256), (256, 256, 128), (128, 128, 128), (128, 128, 128), (128, 128, 128)))
    close(st["voxel_feats"], want, rtol=3e-4, atol=2e-4)          # 34 chained tensor-core layers
    neck, mask = O.voxel2point_neck(pts[:, :5], coors, N(st["voxel_feats"]), inv, cfg["seg_voxel_size"], cfg["point_cloud_range"])
    assert mask.all() and np.array_equal(N(st["pts_lidar_feats"]), neck)


def test_enhance_stage(frame):
    st, sd = frame["st"], frame["sd"]
    scores, ids, cam, fg, ov = OF.img_scores(frame["pts"][:, 5:8], frame["mask"], frame["l2i"], frame["anno"])
    assert np.array_equal(N(st["fg"]).astype(bool), fg) and np.array_equal(N(st["overlap"]), ov) and np.array_equal(N(st["cam_sel"]), cam)
    assert np.array_equal(N(st["img_scores"]), scores)
    assert fg.mean() > 0.05
    img = O.mlp_from_state_dict(scores, sub(sd, "segmentor_updated_mlp."), "ln", "gelu", 1e-3)
    close(st["seg_feats"], (img + N(st["pts_lidar_feats"])).astype(np.float32))
    logits, votes = OM.vote_seg_head(N(st["seg_feats"]), sub(sd, "segmentation_head."))
    close(st["seg_logits"], logits)
    close(st["seg_vote_preds"], votes)
    assert np.array_equal(N(st["offsets"]), O.decode_vote_targets(N(st["seg_vote_preds"])))


def test_frustum_stage(frame):
    st, sd, pts = frame["st"], frame["sd"], frame["pts"]
    ids = O.points_in_mask(pts[:, 5:8], frame["mask"], frame["l2i"])
    rows, obj = OF.frustum_rows(ids)
    assert np.array_equal(N(st["frustum_rows"]), rows)
    sir_coors = np.stack([np.zeros_like(obj), np.zeros_like(obj), obj], 1)
    assert np.array_equal(N(st["frustum_sir_coors"]), sir_coors)
    assert len(rows) > int(N(st["fg"]).sum()) > 0                    # overlap duplication happened
    fgw = OF.point_fg_weights(N(st["seg_logits"]))
    close(st["point_fg_weights"], fgw, atol=1e-6)
    f_cluster, center, ccoors, _ = OF.cluster_delta_weighted(pts[rows, :3], sir_coors, N(st["point_fg_weights"])[rows])
    close(st["frustum_f_cluster"], f_cluster)
    close(st["frustum_obj_centers"], center)
    assert np.array_equal(N(st["frustum_obj_coors"]), ccoors)
    assert np.array_equal(N(st["frustum_pts"]), pts[rows, :5]) and np.array_equal(N(st["frustum_pts_feats"]), N(st["seg_feats"])[rows])
    _, cl, oc = OM.sir(pts[rows, :5], N(st["seg_feats"])[rows], sir_coors, N(st["frustum_f_cluster"]), sub(sd, "frustum_sir."), 3, [20, 20, 4])
    assert np.array_equal(oc, ccoors)
    close(st["frustum_obj_feats"][:, :768], cl)
    preds, enc = OF.encode_preds_2d(frame["anno"], ccoors[:, 2], frame["mask"].shape[-1], frame["mask"].shape[-2], 10)
    assert np.array_equal(N(st["frustum_preds_2d"]), preds) and np.array_equal(N(st["frustum_enc_2d"]), enc)
    img = O.mlp_from_state_dict(enc, sub(sd, "encode_2d_mlp."), "ln", "gelu", 1e-3)
    close(st["frustum_obj_feats"][:, 768:], img)
    cls, reg = head_oracle(N(st["frustum_obj_feats"]), sd, "frustum_obj_head.")
    close(st["frustum_cls"], cls, rtol=2e-4, atol=1e-4)
    close(st["frustum_reg"], reg, rtol=2e-4, atol=1e-4)


def test_fsd_stage(frame):
    st, sd, pts = frame["st"], frame["sd"], frame["pts"]
    cfg = FSFM.NUSC
    data = dict(p=pts[:, :5], l=N(st["seg_logits"]), v=N(st["seg_vote_preds"]), f=N(st["seg_feats"]), o=N(st["offsets"]))
    pre, uniq, _ = OF.pre_voxelize(data, pts[:, :5], cfg["pre_voxelization_size"], cfg["point_cloud_range"])
    assert np.array_equal(N(st["pre_coors"]), uniq)
    for k, name in (("p", "pre_points"), ("l", "pre_logits"), ("v", "pre_votes"), ("f", "pre_feats"), ("o", "pre_offsets")):
        close(st[name], pre[k])
    groups = frame["model"].groups
    score, centers = OF.group_sample(N(st["pre_logits"]), N(st["pre_points"]), N(st["pre_offsets"]), groups, cfg["score_thresh"])
    close(st["group_score"], score, atol=1e-6)
    close(st["group_centers"], centers, atol=1e-5)
    g_score, g_centers = N(st["group_score"]), N(st["group_centers"])
    rows_all, inds_all = [], []
    for g in range(len(groups)):
        idx = np.flatnonzero(g_score[:, g] > np.float32(cfg["score_thresh"][g]))
        if len(idx) == 0:
            idx = np.zeros(1, np.int64)
        labels, keep = OF.cluster_assign_single(g_centers[idx, g], cfg["cluster_voxel_size"][g], cfg["point_cloud_range"],
                                                cfg["connected_dist"][g], cfg["min_points"])
        rows_all.append(idx[keep])
        inds_all.append(np.stack([np.full(len(keep), g), np.zeros(len(keep), np.int64), labels], 1))
    rows_all, inds_all = np.concatenate(rows_all), np.concatenate(inds_all)
    assert np.array_equal(N(st["fsd_rows"]), rows_all)
    assert np.array_equal(N(st["pts_cluster_inds"]), inds_all)          # CCL labels bit-exact through the whole assigner
    assert len(rows_all) > 50
    # extract_feat
    center_preds = N(st["fsd_center_preds"])
    cxyz, ccoors, cinv = O.scatter_v2(center_preds, inds_all, "avg")
    close(st["fsd_obj_centers"], cxyz)
    close(st["fsd_f_cluster"], (N(st["fsd_pts"])[:, :3] - cxyz[cinv]).astype(np.float32))
    want_feats = np.concatenate([N(st["pre_logits"])[rows_all], N(st["pre_votes"])[rows_all], N(st["pre_feats"])[rows_all]], 1)
    assert np.array_equal(N(st["fsd_pts_feats"]), want_feats)
    _, cl, oc = OM.sir(N(st["fsd_pts"]), want_feats, inds_all, N(st["fsd_f_cluster"]), sub(sd, "backbone."), 3, [20, 20, 4])
    assert np.array_equal(N(st["fsd_obj_coors"]), oc)
    close(st["fsd_obj_feats"], cl)
    cls, reg = head_oracle(N(st["fsd_obj_feats"]), sd, "bbox_head.")
    close(st["fsd_cls"], cls, rtol=2e-4, atol=1e-4)
    close(st["fsd_reg"], reg, rtol=2e-4, atol=1e-4)


def test_combine_stage(frame):
    st, sd = frame["st"], frame["sd"]
    fr = O.mlp_from_state_dict(N(st["frustum_obj_feats"]), sub(sd, "combine_frustum_feat_mlp."), "ln", "gelu", 1e-3)
    fs = O.mlp_from_state_dict(N(st["fsd_obj_feats"]), sub(sd, "combine_fsd_feat_mlp."), "ln", "gelu", 1e-3)
    close(st["obj_feats"], np.concatenate([fr, fs], 0), rtol=2e-4, atol=1e-4)
    kf = len(fr)
    fc = N(st["fsd_obj_coors"])
    assert np.array_equal(N(st["obj_coors"])[kf:], np.stack([fc[:, 1], fc[:, 0], fc[:, 2] + 1000], 1))
    assert st["obj_cls"].shape == (kf + len(fs), 10) and st["obj_reg"].shape == (kf + len(fs), 10)


@pytest.mark.parametrize("thr", [0.1, 0.02, 0.9])
def test_group_cluster_equals_loop(cuda, frame, thr):
    """All class groups in one pass (csrc/group_cluster.cu) == the reference's per-group loop (single_stage_fsd.py:822-842,
    936-982), bit for bit: candidate rows, (cls, batch, cluster) ids, centres and everything computed from them.
    thr 0.9 empties groups (the `at least one point` and `keep everything` fallbacks), 0.02 floods them."""
    model = frame["model"]
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda)
    args = (T(frame["pts"]), T(frame["mask"]), T(frame["anno"]), T(frame["l2i"]))
    old = model.cfg
    model.cfg = dict(old, score_thresh=[thr] * 6)
    try:
        with torch.no_grad():
            model.group_loop = True
            a = model(*args)
            model.group_loop = False
            b = model(*args)
    finally:
        model.group_loop, model.cfg = False, old
    assert a["fsd_rows"].numel() > 0
    for k in ("fsd_rows", "pts_cluster_inds", "fsd_center_preds", "fsd_obj_coors", "fsd_obj_centers", "fsd_obj_feats", "fsd_cls"):
        assert torch.equal(a[k], b[k]), k


def test_refine_stage(cuda, frame):
    """Query refinement (FSF.each_stage_refine / query_feat_refine, FSF.py:1009-1083) on the frame's combined queries, step by
    step against the numpy restatement, each step fed the GPU path's inputs (teacher forcing)."""
    model, sd, pts = frame["model"], frame["sd"], frame["pts"]
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda)
    with torch.no_grad():
        st = model(T(pts), T(frame["mask"]), T(frame["anno"]), T(frame["l2i"]))
        st = model.refine(st, T(pts))
    rois = N(st["refine0_rois"])
    want_rois = O.decode_boxes(N(st["obj_reg"]), N(st["obj_centers"]))
    np.testing.assert_allclose(rois, want_rois, rtol=2e-5, atol=2e-6)
    # pooling: same (roi, point) pairs up to points that sit on a face of an enlarged box
    wp, wr, wf, amb = O.dynamic_point_pool(rois[:, 1:8], pts[:, :3], [1.0, 1.0, 1.0], 512, 50000, margin=1e-4)
    inds, roi_inds = N(st["refine0_pts_inds"]), N(st["refine0_roi_inds"])
    assert inds.size > 1
    assert (set(zip(roi_inds.tolist(), inds.tolist())) ^ set(zip(wr.tolist(), wp.tolist()))) <= amb
    # RoI head on the GPU's pooled points
    img_feat = O.mlp_from_state_dict(N(st["img_scores"])[inds], sub(sd, "refine_img_mlp.0."), "ln", "gelu", 1e-3)
    feats = np.concatenate([N(st["seg_feats"])[inds], img_feat], 1)
    want_feat, want_mask = OM.fully_sparse_bbox_head(pts[inds, :5], feats, N(st["refine0_local"]), N(st["refine0_offset"]),
                                                     N(st["refine0_margin"]), roi_inds, rois, sub(sd, "refine_sir_layers.0."), 3)
    np.testing.assert_array_equal(N(st["refine0_mask"]), want_mask)
    close(st["refine0_lidar_feat"], want_feat, atol=1e-4)
    # query update and refined head
    cur = O.mlp_from_state_dict(N(st["refine0_lidar_feat"]), sub(sd, "lidar_img_mlp.0."), "ln", "gelu", 1e-3)
    pos = O.mlp_from_state_dict(rois[:, 1:4], sub(sd, "position_encoder.0."), "ln", "gelu", 1e-3)
    query = O.mlp_from_state_dict((cur + N(st["obj_feats"]) + pos).astype(np.float32), sub(sd, "out_proj.0."), "ln", "gelu", 1e-3)
    scale = float(np.abs(query).max())
    close(st["refine0_query"], query, atol=1e-4 * scale)
    cls, reg = head_oracle(N(st["refine0_query"]), sd, "frustum_refined_head.0.")
    close(st["refine0_cls"], cls, atol=1e-4 * max(1.0, float(np.abs(cls).max())))
    close(st["refine0_reg"], reg, atol=1e-4 * max(1.0, float(np.abs(reg).max())))


def test_group_cluster_reference_golden(cuda):
    """ops.group_sample + ops.group_cluster (all class groups in one pass) against the output of the reference's own
    group_sample + ClusterAssigner (tests/golden/group_cluster.npz): rows and (cls, batch, cluster) ids bit-exact."""
    from fullysparsefusion_b200 import ops
    from fullysparsefusion_b200.fsf import NUSC
    from tests.conftest import load_golden
    g = load_golden("group_cluster")
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda)
    groups = [[NUSC["class_names"].index(n) for n in grp] for grp in NUSC["group_names"]]
    _, score, centers = ops.group_sample(T(g["logits"]), groups, xyz=T(g["points"]), offsets=T(g["offsets"]))
    rows, cls, clu, ctr = ops.group_cluster(score, centers, NUSC["score_thresh"], NUSC["cluster_voxel_size"], NUSC["point_cloud_range"],
                                            NUSC["connected_dist"], NUSC["min_points"])
    assert np.array_equal(N(rows), g["rows"])
    assert np.array_equal(np.stack([N(cls), np.zeros(len(g["rows"]), np.int64), N(clu)], 1), g["cluster_inds"])
    np.testing.assert_allclose(N(ctr), g["center_preds"], rtol=1e-5, atol=1e-5)


def test_detections(cuda, frame):
    """Forward → refine → get_bboxes: the final detections against decode + NMS of the oracle on the GPU's refined outputs."""
    model, pts = frame["model"], frame["pts"]
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda)
    with torch.no_grad():
        st = model.refine(model(T(pts), T(frame["mask"]), T(frame["anno"]), T(frame["l2i"])), T(pts))
        boxes, scores, labels = model.get_bboxes(st, score_thr=0.3, nms_thr=0.35, max_num=500)
    rois = O.decode_boxes(N(st["refine0_reg"]), N(st["refine0_centers"]))
    wb, ws, wl, wr, close = O.multiclass_nms(rois[:, 1:], N(st["refine0_cls"]), 0.3, 0.35, 500, margin=2e-4)
    assert len(wb) > 0
    if not close:   # no deciding IoU within 2e-4 of the threshold: the kept set is exact
        assert np.array_equal(N(st["det_rows"]), wr) and np.array_equal(N(labels), wl)
        np.testing.assert_allclose(N(scores), ws, rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(N(boxes), wb, rtol=1e-5, atol=1e-5)
    else:
        assert abs(len(N(labels)) - len(wl)) <= 2


def test_simple_test_entry(cuda, frame):
    """FSF.simple_test (the reference's test entry and argument conventions) = forward → refine → get_bboxes on each sample."""
    model, pts = frame["model"], frame["pts"]
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda)
    metas = [dict(lidar2img=[m for m in frame["l2i"].astype(np.float64)])]           # python-side 4x4 matrices, as the dataset gives
    with torch.no_grad():
        res = model.simple_test([T(pts)], metas, T(frame["mask"])[None], T(frame["anno"])[None], score_thr=0.3)
        st = model.refine(model(T(pts), T(frame["mask"]), T(frame["anno"]), T(frame["l2i"])), T(pts))
        boxes, scores, labels = model.get_bboxes(st, score_thr=0.3, nms_thr=0.35, max_num=500)
    assert len(res) == 1 and set(res[0]) == {"boxes_3d", "scores_3d", "labels_3d"}
    assert not res[0]["boxes_3d"].is_cuda and res[0]["boxes_3d"].shape[1] == 9
    assert len(boxes) > 0 and abs(len(res[0]["labels_3d"]) - len(labels)) <= 2      # two passes over the same frame
    if len(res[0]["labels_3d"]) == len(labels):
        assert torch.equal(res[0]["labels_3d"], labels.cpu())
        torch.testing.assert_close(res[0]["boxes_3d"], boxes.cpu(), rtol=1e-4, atol=1e-4)
        torch.testing.assert_close(res[0]["scores_3d"], scores.cpu(), rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("points,sweeps", [(34000, 1), (300000, 10)])
def test_full_frame_against_cpu_port_at_benchmark_sizes(cuda, points, sweeps):
    """BASELINE configs[1] (single sweep, 34 k points) and configs[2] (10 sweeps, 300 k points: the benchmarked shape) at the real
    1600 x 900 image size, every stage of FSF.simple_test including refinement and the final boxes, against the torch-CPU port
    on the same frame and weights (the numpy oracle is too slow here; the port itself is pinned to it in
    tests/test_cpu_port_vs_oracle.py).  Same comparison bench.py prints as `parity`."""
    import copy

    import bench
    from oracle import fsf_torch_cpu as P

    frame = bench.synth_frame(points, sweeps, seed=3)
    dev_frame = {k: v.to(cuda) for k, v in frame.items()}
    model = bench.make_model().to(cuda)
    with torch.no_grad():
        st0 = model(dev_frame["points"], dev_frame["mask"], dev_frame["anno"], dev_frame["lidar2img"])
        bench.calibrate_seg_head(model, st0["seg_logits"])
        stages, st = model.stages(dev_frame["points"], dev_frame["mask"], dev_frame["anno"], dev_frame["lidar2img"])
        for _, fn in stages:
            fn()
        model.refine(st, dev_frame["points"])
        model.get_bboxes(st)
        torch.cuda.synchronize()
        cpu_model = copy.deepcopy(model).cpu()
        cpu = P.CpuFSF(cpu_model)
        cstages, cst = cpu.stages(frame["points"], frame["mask"], frame["anno"], frame["lidar2img"])
        for _, fn in cstages + cpu.extra_stages:
            fn()
    par = bench.frame_parity(st, cst)
    assert not par["breach"], par
    assert par["rel_p999_by_tensor"]["seg_logits"] < 5e-4 and par["rel_p999_by_tensor"]["voxel_feats"] < 5e-4, par
    assert par["counts_gpu_cpu"]["voxels"][0] == par["counts_gpu_cpu"]["voxels"][1], par                     # the same voxel set
    assert par["detections_matched"] is None or par["detections_matched"] >= 0.97, par


def test_scatter_plan_on_empty_input(cuda):
    """torch.unique survives empty input (an empty frame / every point filtered out); so does the ranking path."""
    from fullysparsefusion_b200 import modules as M, ops
    empty = torch.zeros((0, 4), dtype=torch.int32, device=cuda)
    u, inv, cnt = ops.unique_rows(empty, lo=[0, 0, 0, 0], ext=[1, 40, 512, 512], return_counts=True)
    assert u.shape == (0, 4) and inv.numel() == 0 and cnt.numel() == 0
    res = ops.unique_rows(empty, lo=[0, 0, 0, 0], ext=[1, 40, 512, 512], return_unique=False, return_index=True)
    assert res[0] is None and res[3].m == 0
    plan = M.ScatterPlan(empty, lo=[0, 0, 0, 0], ext=[1, 40, 512, 512], want_index=True)
    assert plan.m == 0


def test_av2_frame_against_cpu_port(cuda):
    """BASELINE configs[3]: the stock Argoverse 2 configuration (FSF_AV2_config.py) — ~107 k 4-d points, seven ring cameras with ONE
    int32 id plane each, 26 classes in six groups, the 32 x 2048 x 2048 grid, a four-stage 64-channel U-Net, the 32-channel per-point
    encoding of the selected 2-D object (is_argo) and no velocity head — every stage of simple_test against the torch-CPU port."""
    import copy

    import bench
    from oracle import fsf_torch_cpu as P

    frame = bench.synth_frame(107000, 1, seed=4, config="av2")
    dev_frame = {k: v.to(cuda) for k, v in frame.items()}
    model = bench.make_model(config="av2").to(cuda)
    with torch.no_grad():
        st0 = model(dev_frame["points"], dev_frame["mask"], dev_frame["anno"], dev_frame["lidar2img"])
        bench.calibrate_seg_head(model, st0["seg_logits"])
        stages, st = model.stages(dev_frame["points"], dev_frame["mask"], dev_frame["anno"], dev_frame["lidar2img"])
        for _, fn in stages:
            fn()
        model.refine(st, dev_frame["points"])
        model.get_bboxes(st)
        torch.cuda.synchronize()
        cpu = P.CpuFSF(copy.deepcopy(model).cpu())
        cstages, cst = cpu.stages(frame["points"], frame["mask"], frame["anno"], frame["lidar2img"])
        for _, fn in cstages + cpu.extra_stages:
            fn()
    assert st["seg_logits"].shape[1] == 27 and st["seg_feats"].shape[1] == 67 and st["img_scores"].shape[1] == 32
    assert st["refine0_reg"].shape[1] == 8 and st["det_boxes"].shape[1] == 7          # no velocity
    par = bench.frame_parity(st, cst)
    assert not par["breach"], par
    assert par["counts_gpu_cpu"]["voxels"][0] == par["counts_gpu_cpu"]["voxels"][1], par
    assert par["rel_p999_by_tensor"]["seg_logits"] < 5e-4, par
