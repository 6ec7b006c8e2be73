"""FSF inference forward (one frame per call) on the B200 hot path.

Mirrors projects/mmdet3d_plugin/models/detectors/FSF.py::simple_test (:1114-1176) up to and including
combine_frustum_and_fsd (:657-692), with the segmentor of single_stage_fsd.py (VoteSegmentor.extract_feat
:228-245) inlined:

  segment      voxelize → DynamicScatterVFE → SimpleSparseUNet → Voxel2PointScatterNeck
  enhance      img_cross_attn (projection ⊕ sampling ⊕ camera select ⊕ score lookup → MLP) + VoteSegHead
  frustum      camera queries: fg extraction, overlap duplication, weighted centroids, SIR, FrustumClusterHead
  fsd          LiDAR queries: pre_voxelize, group_sample, ClusterAssigner (CCL), SIR, SparseClusterHeadV2
  combine      the two 1024-wide fusion MLPs

`FSF.refine` continues with the query-refinement stage (decode_stage_bboxes → dynamic point pooling →
FullySparseBboxHead → query MLPs → refined head) and `FSF.get_bboxes` with the final box decode and
rotated multi-class NMS (rank 3).  One sample per call (samples_per_gpu = 1 in both stock configs).
All arithmetic goes through the C-ABI ops; torch is used for allocation, views and concatenation only.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Tuple

import torch
import torch.nn.functional as F
from torch import nn

from . import modules as M
from . import ops

# ---- the stock nuScenes configuration (projects/configs/nuScenes/FSF_nuScenes_config.py) ----------------
NUSC = dict(
    class_names=["car", "truck", "trailer", "bus", "construction_vehicle", "bicycle", "motorcycle", "pedestrian",
                 "traffic_cone", "barrier"],
    group_names=[["car"], ["truck", "construction_vehicle"], ["bus", "trailer"], ["barrier"], ["motorcycle", "bicycle"],
                 ["pedestrian", "traffic_cone"]],
    seg_voxel_size=(0.2, 0.2, 0.2), point_cloud_range=[-51.2, -51.2, -5.0, 51.2, 51.2, 3.0], sparse_shape=[40, 512, 512],
    score_thresh=[0.1] * 6, pre_voxelization_size=(0.1, 0.1, 0.1),
    cluster_voxel_size=[(0.3, 0.3, 8), (0.3, 0.3, 8), (0.3, 0.3, 8), (0.1, 0.1, 8), (0.2, 0.2, 8), (0.05, 0.05, 8)],
    connected_dist=[0.6, 0.6, 0.6, 0.2, 0.4, 0.1], min_points=2, num_cams=6,
)
NUSC.update(
    point_dim=5, is_argo=False, code_size=10,
    unet=dict(base_channels=64, output_channels=128,
              encoder_channels=((128,), (128, 128, 128), (128, 128, 128), (256, 256, 256), (512, 512, 512)),
              encoder_paddings=((1,), (1, 1, 1), (1, 1, 1), ((0, 1, 1), 1, 1), (1, 1, 1)),
              decoder_channels=((512, 512, 256), (256, 256, 128), (128, 128, 128), (128, 128, 128), (128, 128, 128)),
              decoder_paddings=((1, 1), (1, 0), (1, 0), (0, 0), (0, 1))))

# ---- the stock Argoverse 2 configuration (projects/configs/Argoverse2/FSF_AV2_config.py): 4-d points, the +-204.8 m grid
# [32, 2048, 2048] (:11,86), a four-stage 64-channel U-Net (:83-95), 26 classes in six groups (:12-49), seven ring cameras
# with ONE int32 id plane each (datasets/pipelines/loading.py:169-186), per-point 2-D encoding of the selected object
# (is_argo: bbox / score / one-hot category = 32 channels, FSF.py:449-474,537-552), no velocity (code_size 8, :164,259)
_AV2_CLASSES = ["Regular_vehicle", "Pedestrian", "Bicyclist", "Motorcyclist", "Wheeled_rider", "Bollard", "Construction_cone", "Sign",
                "Construction_barrel", "Stop_sign", "Mobile_pedestrian_crossing_sign", "Large_vehicle", "Bus", "Box_truck", "Truck",
                "Vehicular_trailer", "Truck_cab", "School_bus", "Articulated_bus", "Message_board_trailer", "Bicycle", "Motorcycle",
                "Wheeled_device", "Wheelchair", "Stroller", "Dog"]
AV2 = dict(
    class_names=_AV2_CLASSES,
    group_names=[_AV2_CLASSES[:1], _AV2_CLASSES[1:5], _AV2_CLASSES[5:11], _AV2_CLASSES[11:20], _AV2_CLASSES[20:25], _AV2_CLASSES[25:]],
    seg_voxel_size=(0.2, 0.2, 0.2), point_cloud_range=[-204.8, -204.8, -3.2, 204.8, 204.8, 3.2], sparse_shape=[32, 2048, 2048],
    score_thresh=[0.4, 0.25, 0.25, 0.25, 0.25, 0.25], pre_voxelization_size=(0.1, 0.1, 0.1),
    cluster_voxel_size=[(0.3, 0.3, 6.4), (0.05, 0.05, 6.4), (0.08, 0.08, 6.4), (0.5, 0.5, 6.4), (0.1, 0.1, 6.4), (0.08, 0.08, 6.4)],
    connected_dist=[0.6, 0.1, 0.15, 1.0, 0.2, 0.15], min_points=2, num_cams=7,
    point_dim=4, is_argo=True, code_size=8,
    unet=dict(base_channels=64, output_channels=64,
              encoder_channels=((64,), (64, 64, 64), (64, 64, 64), (128, 128, 128)),
              encoder_paddings=((1,), (1, 1, 1), (1, 1, 1), ((0, 1, 1), 1, 1)),
              decoder_channels=((128, 128, 64), (64, 64, 64), (64, 64, 64), (64, 64, 64)),
              decoder_paddings=((1, 0), (1, 0), (0, 0), (0, 1))))
CONFIGS = {"nuscenes": NUSC, "av2": AV2}

BN = dict(type="naiveSyncBN1d", eps=1e-3, momentum=0.01)
LN3 = dict(type="LN", eps=1e-3)
LN5 = dict(type="LN")


class FSDSeparateHead(nn.Module):
    """models/dense_heads/sparse_cluster_head_v2.py:17-41."""

    def __init__(self, in_channels, attrs, norm_cfg=LN5, act="relu"):
        super().__init__()
        self.attrs = attrs
        for name, (out_dim, num_layer, hidden) in attrs.items():
            setattr(self, name, M.build_mlp(in_channels, [hidden] * num_layer + [out_dim], norm_cfg, is_head=True, act=act))

    def forward(self, x):
        return {name: getattr(self, name)(x) for name in self.attrs}


class SparseClusterHeadV2(nn.Module):
    """SparseClusterHead.__init__ shared MLP (sparse_cluster_head.py:75-77) + SparseClusterHeadV2.forward
    (sparse_cluster_head_v2.py:134-168); FrustumClusterHead inherits this forward."""

    def __init__(self, num_classes, in_channel, shared_mlp_dims, tasks, common_attrs, num_cls_layer, cls_hidden_dim,
                 separate_head, norm_cfg=LN5, act="relu", bbox_coder=None, class_names=None, train_cfg=None, test_cfg=None,
                 shared_dropout=0, **training_only):
        # `training_only` swallows what the stock configs also pass and inference never reads: loss_cls / loss_center / loss_size /
        # loss_rot / loss_vel / loss_iou, cls_mlp / reg_mlp / iou_mlp (None in the configs), corner_loss_cfg, enlarge_width, as_rpn,
        # init_cfg, and FrustumClusterHead's assigner / num_objs / vis_dir / use_one_to_one (sparse_cluster_head_v2.py:47-76,
        # frustum_cluster_head.py:21-54)
        super().__init__()
        assert not shared_dropout, "dropout is a training-time option"
        self.bbox_coder_cfg, self.class_names, self.train_cfg, self.test_cfg = bbox_coder, class_names, train_cfg, test_cfg
        self.shared_mlp = M.build_mlp(in_channel, list(shared_mlp_dims), norm_cfg, act=act) if len(shared_mlp_dims) else None
        sep_in = shared_mlp_dims[-1] if len(shared_mlp_dims) else in_channel
        self.task_heads = nn.ModuleList()
        for t in tasks:
            attrs = dict(common_attrs)
            attrs["score"] = (len(t["class_names"]), num_cls_layer, cls_hidden_dim)
            self.task_heads.append(FSDSeparateHead(sep_in, attrs, norm_cfg=separate_head.get("norm_cfg", LN5),
                                                   act=separate_head.get("act", "relu")))

    def forward(self, feats, pts_xyz=None, pts_inds=None):
        if self.shared_mlp is not None:
            feats = self.shared_mlp(feats)
        cls_list, reg_list = [], []
        for h in self.task_heads:
            r = h(feats)
            cls_list.append(r["score"])
            parts = [r["center"], r["dim"], r["rot"]] + ([r["vel"]] if "vel" in r else [])
            reg_list.append(torch.cat(parts, dim=-1))
        return dict(cls_logits=cls_list, reg_preds=reg_list)


def _head(in_channel, class_names, code_size=10):
    attrs = dict(center=(3, 2, 128), dim=(3, 2, 128), rot=(2, 2, 128))
    if code_size == 10:   # nuScenes regresses a velocity; the AV2 config (code_size 8) has no `vel` attribute
        attrs["vel"] = (2, 2, 128)
    return SparseClusterHeadV2(
        num_classes=len(class_names), in_channel=in_channel, shared_mlp_dims=[1024, 1024],
        tasks=[dict(num_class=len(class_names), class_names=class_names)],
        common_attrs=attrs, num_cls_layer=2,
        cls_hidden_dim=128, separate_head=dict(norm_cfg=LN5, act="gelu"), norm_cfg=LN5, act="relu")


class FSF(nn.Module):
    """The FSF detector's inference forward on synthetic or real frames (random init unless weights are loaded)."""

    def __init__(self, cfg=NUSC):
        """cfg: one of CONFIGS' dicts (or its name, "nuscenes" / "av2"): every width below is derived from it the way the stock
        config files spell them out (point dims P, classes nc, U-Net output → point feature width F = out + 3)."""
        super().__init__()
        if isinstance(cfg, str):
            cfg = CONFIGS[cfg]
        self.cfg = cfg
        nc = len(cfg["class_names"])
        self.num_classes = nc
        self.is_argo = bool(cfg.get("is_argo", False))
        P = self.point_dim = int(cfg.get("point_dim", 5))
        code = int(cfg.get("code_size", 10))
        unet = cfg.get("unet", NUSC["unet"])
        F_pt = unet["output_channels"] + 3            # neck output: voxel feature ‖ xyz - voxel centre (131 / 67)
        enc2d = 5 + nc + 1                             # bbox (4) + score + one-hot category incl. "none" (16 / 32)
        pt2d = enc2d if self.is_argo else nc           # per-point image input: the 32-channel encoding (AV2) or nc class scores
        self.groups = [[cfg["class_names"].index(n) for n in g] for g in cfg["group_names"]]
        # segmentor (FSF_nuScenes_config.py:33-102; FSF_AV2_config.py:60-132)
        self.voxel_encoder = M.DynamicScatterVFE(in_channels=P, feat_channels=[64, 64], voxel_size=cfg["seg_voxel_size"],
                                                 with_cluster_center=True, with_voxel_center=True,
                                                 point_cloud_range=cfg["point_cloud_range"], norm_cfg=BN, unique_once=True)
        self.backbone_unet = M.SimpleSparseUNet(in_channels=64, sparse_shape=cfg["sparse_shape"], norm_cfg=BN, **unet)
        self.decode_neck = M.Voxel2PointScatterNeck(voxel_size=cfg["seg_voxel_size"], point_cloud_range=cfg["point_cloud_range"])
        self.segmentation_head = M.VoteSegHead(in_channel=F_pt, hidden_dims=[128, 128], num_classes=nc,
                                               norm_cfg=dict(type="naiveSyncBN1d"), act_cfg=dict(type="ReLU"))
        # FSF.__init__ (FSF.py:100-164)
        self.segmentor_updated_mlp = M.build_mlp(pt2d, [128, F_pt], LN3, is_head=True, act="gelu")
        nn.init.constant_(self.segmentor_updated_mlp[-1].weight, 0.0)   # FSF.py:142-143
        nn.init.constant_(self.segmentor_updated_mlp[-1].bias, 0.0)
        self.encode_2d_mlp = M.build_mlp(enc2d, [128, 128], LN3, is_head=False, act="gelu")
        sir_kw = dict(num_blocks=3, feat_channels=[[128, 128]] * 3, rel_mlp_hidden_dims=[[16, 32]] * 3, norm_cfg=LN3, mode="max",
                      xyz_normalizer=[20, 20, 4], act="gelu", unique_once=True)
        fsd_feat = (nc + 1) + 3 * (nc + 1) + F_pt      # pre-voxel logits ‖ votes ‖ features (175 on nuScenes)
        self.fsd_feat_dims = (nc + 1, 3 * (nc + 1), F_pt)
        self.backbone = M.SIR(in_channels=[P + fsd_feat, P + 128, P + 128], **sir_kw)       # LiDAR queries (:113-124; AV2 :150-161)
        self.frustum_sir = M.SIR(in_channels=[P + F_pt, P + 128, P + 128], **sir_kw)          # camera queries (:201-212; AV2 :241-252)
        self.bbox_head = _head(128 * 3 * 2, cfg["class_names"], code)
        self.frustum_obj_head = _head(128 * 3 * 2 + 128, cfg["class_names"], code)
        self.combine_frustum_feat_mlp = M.build_mlp(128 * 3 * 2 + 128, [1024], LN3, act="gelu")
        self.combine_fsd_feat_mlp = M.build_mlp(128 * 3 * 2, [1024], LN3, act="gelu")
        # query refinement (FSF.__init__ :144-164; configs :275-404 / AV2 :318-426): one extra stage in the stock configs
        self.num_extra_stages = int(cfg.get("num_extra_stages", 1))
        self.roi_extractor = M.DynamicPointROIExtractor(extra_wlh=[1.0, 1.0, 1.0], max_inbox_point=512, debug=False)
        embed = 1024
        self.refine_sir_layers = nn.ModuleList([M.FullySparseBboxHead(
            num_classes=nc, num_blocks=3, in_channels=[P + F_pt + 32 + 13, P + 128 + 13, P + 128 + 13], feat_channels=[[128, 128]] * 3,
            rel_mlp_hidden_dims=[[16, 32]] * 3, rel_mlp_in_channels=[13] * 3, xyz_normalizer=[20, 20, 4], act="gelu", geo_input=True,
            use_middle_cluster_feature=True, norm_cfg=LN3, unique_once=True) for _ in range(self.num_extra_stages)])
        self.refine_img_mlp = nn.ModuleList([M.build_mlp(pt2d, [32, 32], LN3, is_head=False, act="gelu") for _ in range(self.num_extra_stages)])
        self.lidar_img_mlp = nn.ModuleList([M.build_mlp(128 * 3 * 2, [embed, embed], LN3, act="gelu") for _ in range(self.num_extra_stages)])
        self.position_encoder = nn.ModuleList([M.build_mlp(3, [embed, embed], LN3, act="gelu") for _ in range(self.num_extra_stages)])
        self.out_proj = nn.ModuleList([M.build_mlp(embed, [embed, embed], LN3, act="gelu", is_head=True) for _ in range(self.num_extra_stages)])
        self.frustum_refined_head = nn.ModuleList([_head(embed, cfg["class_names"], code) for _ in range(self.num_extra_stages)])
        self.fsd_begin_idx = 1000
        self.group_loop = False   # True: the reference's per-group Python loop (kept for the equivalence test)
        self.eval()

    def refresh(self):
        for m in self.modules():
            if m is not self and hasattr(m, "refresh"):
                m.refresh()

    def _cluster_groups_loop(self, score, centers, dev):
        """group_sample's selection + ClusterAssigner.forward_single_class group by group, as the reference loops
        (single_stage_fsd.py:822-842, 936-982)."""
        cfg = self.cfg
        rng = cfg["point_cloud_range"]
        sel_rows, cls_ids, clu_ids, ctr_list = [], [], [], []
        for g in range(len(self.groups)):
            idx = ops.compact_indices(ops.threshold_mask(score, g, cfg["score_thresh"][g]))
            if idx.numel() == 0:                                                     # at least one point per sample (:833-835)
                idx = torch.zeros(1, dtype=torch.int32, device=dev)
            ctr = ops.gather_rows(centers.view(-1, 3 * len(self.groups))[:, 3 * g:3 * g + 3].contiguous(), idx)
            cv = ops.voxelize(ctr, cfg["cluster_voxel_size"][g], rng, floor_mode=1, order_xyz=True, check_range=False)
            cc4 = F.pad(cv, (1, 0), value=0)
            _, inv, cnt = ops.unique_rows(cc4, return_counts=True, return_unique=False, inv_dtype=torch.int32)
            keep = ops.compact_indices(ops.count_mask(cnt, inv, cfg["min_points"]))
            if keep.numel() == 0:                                                    # `valid_mask = ~valid_mask` (:953-955)
                keep = torch.arange(idx.numel(), dtype=torch.int32, device=dev)
            ctr_k = ops.gather_rows(ctr, keep)
            plan_c = M.ScatterPlan(ops.gather_int_rows(cc4, keep))
            sampled_centers = plan_c.reduce(ctr_k, "mean")
            labels = ops.connected_components(sampled_centers, None, cfg["connected_dist"][g])   # single-batch variant (:977)
            clu = ops.gather_int_rows(labels.view(-1, 1), plan_c.inv32)
            sel_rows.append(ops.gather_int_rows(idx.view(-1, 1), keep).view(-1))
            cls_ids.append(torch.full((keep.numel(), 1), g, dtype=torch.int32, device=dev))
            clu_ids.append(clu)
            ctr_list.append(ctr_k)
        rows = torch.cat(sel_rows)
        pts_cluster_inds = torch.cat([torch.cat(cls_ids), torch.zeros((rows.numel(), 1), dtype=torch.int32, device=dev),
                                      torch.cat(clu_ids)], dim=1)                    # (cls, batch, cluster) (:145-152)
        return rows, pts_cluster_inds, torch.cat(ctr_list)

    # ------------------------------------------------------------------------------------------------
    def stages(self, points: torch.Tensor, mask_data: torch.Tensor, mask_anno: torch.Tensor, lidar2img: torch.Tensor
               ) -> Tuple[List[Tuple[str, Callable[[], None]]], Dict[str, torch.Tensor]]:
        """points [N,P+3] f32 (nuScenes: x,y,z,intensity,dt; AV2: x,y,z,intensity; + un-augmented xyz), mask_data
        [cams,classes,H,W] u8 (nuScenes: 10 class planes) or i32 (AV2: one plane), mask_anno f32 [250,9], lidar2img f32 [cams,4,4].
        Returns ([(stage name, fn)], state)."""
        st: Dict[str, torch.Tensor] = {}
        cfg = self.cfg
        dev = points.device
        P = self.point_dim
        noaug = slice(P, P + 3)
        rng, vs = cfg["point_cloud_range"], cfg["seg_voxel_size"]
        grid_zyx = cfg["sparse_shape"]

        def segment():
            pts5 = points[:, :P]
            coors3 = ops.voxelize(points, vs, rng, floor_mode=0)                       # single_stage_fsd.py:217-219
            coors4 = F.pad(coors3, (1, 0), value=0)                                    # batch pad (:222-225)
            plan = M.ScatterPlan(coors4, lo=[0, 0, 0, 0], ext=[1] + list(grid_zyx), want_index=True)
            voxel_feats, voxel_coors, _ = self.voxel_encoder(pts5, coors4, return_inv=True, plan=plan)
            rb, _ = self.backbone_unet.build_rulebooks(voxel_coors, plan.index, 1)
            x = self.backbone_unet(dict(voxel_feats=voxel_feats, voxel_coors=voxel_coors), rulebooks=rb)[0]["voxel_feats"]
            neck_out, mask = self.decode_neck(pts5, coors4, x, plan.inv32, -1)
            st.update(coors4=coors4, voxel_coors=voxel_coors, vfe_feats=voxel_feats, voxel_feats=x, pts_lidar_feats=neck_out,
                      valid_pts_mask=mask, voxel2point_inds=plan.inv32)

        def enhance():
            # img_cross_attn (FSF.py:694-728): the whole gather/select/lookup chain is one kernel
            if self.is_argo:   # one id plane: the selected object's annotation row → bbox / score / one-hot category (FSF.py:449-474,537-552)
                ids, cam, fg, ov = ops.project_sample_select(points[:, noaug], lidar2img, mask_data, want_overlap=True, want_ids=True)
                _, scores = ops.encode_preds_2d(mask_anno, ids, mask_data.shape[-1], mask_data.shape[-2], self.num_classes, coor_col=0)
            else:
                _, cam, fg, ov, scores = ops.project_sample_select(points[:, noaug], lidar2img, mask_data, want_overlap=True,
                                                                   anno=mask_anno, anno_col=4, want_ids=False)
            img_feat = self.segmentor_updated_mlp(scores)
            st["img_scores"] = scores
            pts_feats = ops.add_(img_feat, st["pts_lidar_feats"])                      # :790
            logits, vote_preds = self.segmentation_head(pts_feats)
            offsets = self.segmentation_head.decode_vote_targets(vote_preds)
            st.update(fg=fg, overlap=ov, cam_sel=cam, seg_feats=pts_feats, seg_logits=logits, seg_vote_preds=vote_preds,
                      offsets=offsets)

        def frustum():
            pts5 = points[:, :P]
            fgw, _, _ = ops.group_sample(st["seg_logits"], want_fg_weight=True)         # get_point_fg_weights (:345-355)
            rows, sir_coors, n_fg = ops.frustum_rows(points[:, noaug], lidar2img, mask_data, st["fg"], st["overlap"])
            if rows.numel() == 0:   # fake one object (FSF.py:407-414)
                r_pts = torch.zeros((1, P), device=dev)
                r_feat = torch.zeros((1, st["seg_feats"].size(1)), device=dev)
                sir_coors = torch.zeros((1, 3), dtype=torch.int32, device=dev)
                f_cluster = torch.zeros((1, 3), device=dev)
                center = torch.zeros((1, 3), device=dev)
                plan = M.ScatterPlan(sir_coors)
            else:
                r_pts = ops.gather_rows(pts5, rows)
                r_feat = ops.gather_rows(st["seg_feats"], rows)
                plan = M.ScatterPlan(sir_coors)
                w4 = ops.weighted_xyz(pts5, fgw, rows)                                  # :313-318
                mean4 = plan.reduce(w4, "mean")
                f_cluster, center = ops.cluster_delta(pts5, mean4, plan.inv32, rows)    # :324-329
            _, cluster_feats, out_coors = self.frustum_sir(r_pts, r_feat, sir_coors, f_cluster, plan=plan)
            preds_2d, enc = ops.encode_preds_2d(mask_anno, out_coors, mask_data.shape[-1], mask_data.shape[-2], self.num_classes)
            img_feat = self.encode_2d_mlp(enc)
            obj_feat = torch.cat([cluster_feats, img_feat], dim=-1)
            res = self.frustum_obj_head(obj_feat)
            st.update(frustum_sir_coors=sir_coors, frustum_f_cluster=f_cluster, frustum_pts=r_pts, frustum_pts_feats=r_feat,
                      frustum_enc_2d=enc)
            st.update(frustum_rows=rows, frustum_obj_feats=obj_feat, frustum_obj_centers=center, frustum_obj_coors=out_coors,
                      frustum_cls=res["cls_logits"][0], frustum_reg=res["reg_preds"][0], frustum_preds_2d=preds_2d,
                      point_fg_weights=fgw)

        def fsd():
            pts5 = points[:, :P].contiguous()
            # pre_voxelize (single_stage_fsd.py:585-605)
            c3 = ops.voxelize(pts5, cfg["pre_voxelization_size"], rng, floor_mode=1)
            c4 = F.pad(c3, (1, 0), value=0)
            plan = M.ScatterPlan(c4, lo=[0, 0, 0, 0], ext=[1] + [g * 2 for g in grid_zyx])
            v_pts = plan.reduce(pts5, "mean")
            v_logits = plan.reduce(st["seg_logits"], "mean")
            v_votes = plan.reduce(st["seg_vote_preds"], "mean")
            v_feats = plan.reduce(st["seg_feats"], "mean")
            v_off = plan.reduce(st["offsets"], "mean")
            # group_sample (:802-865)
            _, score, centers = ops.group_sample(v_logits, self.groups, xyz=v_pts, offsets=v_off)
            if self.group_loop:
                rows, pts_cluster_inds, center_preds = self._cluster_groups_loop(score, centers, dev)
            else:
                # every class group in one pass (csrc/group_cluster.cu): same candidates, order and cluster ids as the loop
                rows, cls, clu, center_preds = ops.group_cluster(
                    score, centers, cfg["score_thresh"], cfg["cluster_voxel_size"], rng, cfg["connected_dist"], cfg["min_points"])
                pts_cluster_inds = torch.stack([cls, torch.zeros_like(cls), clu], dim=1)   # (cls, batch, cluster) (:145-152)
            n = rows.numel()
            s_pts = ops.gather_rows(v_pts, rows)
            d0, d1, d2 = self.fsd_feat_dims                                           # logits ‖ votes ‖ point features
            pts_feats = torch.empty((n, d0 + d1 + d2), dtype=torch.float32, device=dev)
            ops.gather_rows(v_logits, rows, out=pts_feats[:, :d0])
            ops.gather_rows(v_votes, rows, out=pts_feats[:, d0:d0 + d1])
            ops.gather_rows(v_feats, rows, out=pts_feats[:, d0 + d1:])
            # extract_feat (:458-474)
            plan_q = M.ScatterPlan(pts_cluster_inds)
            cluster_xyz_mean = plan_q.reduce(center_preds, "mean")
            f_cluster, cluster_xyz = ops.cluster_delta(s_pts, cluster_xyz_mean, plan_q.inv32)
            _, cluster_feats, cluster_inds = self.backbone(s_pts, pts_feats, pts_cluster_inds, f_cluster, plan=plan_q)
            res = self.bbox_head(cluster_feats)
            st.update(pre_points=v_pts, pre_logits=v_logits, pre_votes=v_votes, pre_feats=v_feats, pre_offsets=v_off,
                      group_score=score, group_centers=centers, fsd_pts=s_pts, fsd_pts_feats=pts_feats,
                      fsd_center_preds=center_preds, fsd_f_cluster=f_cluster)
            st.update(pre_coors=plan.new_coors, fsd_rows=rows, pts_cluster_inds=pts_cluster_inds, fsd_obj_feats=cluster_feats,
                      fsd_obj_centers=cluster_xyz, fsd_obj_coors=cluster_inds, fsd_cls=res["cls_logits"][0],
                      fsd_reg=res["reg_preds"][0])

        def combine():
            fr = self.combine_frustum_feat_mlp(st["frustum_obj_feats"])                  # FSF.py:683-684
            fs = self.combine_fsd_feat_mlp(st["fsd_obj_feats"])
            fc = st["fsd_obj_coors"]
            fsd_re = torch.stack([fc[:, 1], fc[:, 0], fc[:, 2] + self.fsd_begin_idx], dim=1)
            st.update(obj_feats=torch.cat([fr, fs], dim=0),
                      obj_centers=torch.cat([st["frustum_obj_centers"], st["fsd_obj_centers"]], dim=0),
                      obj_coors=torch.cat([st["frustum_obj_coors"], fsd_re], dim=0),
                      obj_cls=torch.cat([st["frustum_cls"], st["fsd_cls"]], dim=0),
                      obj_reg=torch.cat([st["frustum_reg"], st["fsd_reg"]], dim=0))

        return [("segment", segment), ("enhance", enhance), ("frustum", frustum), ("fsd", fsd), ("combine", combine)], st

    @torch.no_grad()
    def refine(self, st: Dict[str, torch.Tensor], points: torch.Tensor) -> Dict[str, torch.Tensor]:
        """Query refinement on the state a forward pass left behind: FSF.multi_stage_refine_test / each_stage_refine /
        query_feat_refine (models/detectors/FSF.py:960-1083) up to the refined heads' logits and regressions (box decoding of
        the final stage and NMS are SURVEY.md section 8f rank 3).  Not part of `stages()`: the benchmarked scope and its CPU
        port end at combine_frustum_and_fsd."""
        pts5 = points[:, :self.point_dim].contiguous()
        obj_centers, obj_reg, res_query_feat = st["obj_centers"], st["obj_reg"], st["obj_feats"]
        out = {}
        for i in range(self.num_extra_stages):
            rois = ops.decode_boxes(obj_reg, obj_centers)                               # decode_stage_bboxes (:1085-1095)
            obj_centers = rois[:, 1:4].contiguous()
            inds, roi_inds, info = self.roi_extractor(pts5, None, rois[:, :8].contiguous())   # :1020-1024
            fake = inds.numel() == 1 and int(inds[0]) < 0                               # nothing pooled: upstream's fake row
            rows = (torch.full((1,), pts5.size(0) - 1, dtype=torch.int32, device=points.device) if fake   # points[-1], as upstream indexes
                    else inds.to(torch.int32))
            ex_pts = ops.gather_rows(pts5, rows)
            ex_feats = ops.gather_rows(st["seg_feats"], rows)
            img_feat = self.refine_img_mlp[i](ops.gather_rows(st["img_scores"], rows))   # img_cross_attn on the pooled points (:1029-1036)
            feats = torch.cat([ex_feats, img_feat], dim=-1)
            lidar_feat, lidar_mask = self.refine_sir_layers[i](ex_pts, feats, info, roi_inds, rois)
            cur = self.lidar_img_mlp[i](lidar_feat)
            pos = self.position_encoder[i](obj_centers)
            query = self.out_proj[i](ops.add_(ops.add_(cur, res_query_feat), pos))     # :1077-1079
            res = self.frustum_refined_head[i](query)
            obj_reg, res_query_feat = res["reg_preds"][0], query
            out.update({f"refine{i}_rois": rois, f"refine{i}_pts_inds": inds, f"refine{i}_roi_inds": roi_inds,
                        f"refine{i}_lidar_feat": lidar_feat, f"refine{i}_mask": lidar_mask, f"refine{i}_query": query,
                        f"refine{i}_cls": res["cls_logits"][0], f"refine{i}_reg": obj_reg, f"refine{i}_centers": obj_centers,
                        f"refine{i}_local": info["local_xyz"], f"refine{i}_offset": info["boundary_offset"],
                        f"refine{i}_margin": info["is_in_margin"]})
        st.update(out)
        return st

    @torch.no_grad()
    def get_bboxes(self, st: Dict[str, torch.Tensor], score_thr: float = 0.01, nms_thr: float = 0.35, max_num: int = 500):
        """Detections of the last refine stage: FrustumClusterHead._get_bboxes_single (dense_heads/frustum_cluster_head.py:
        595-698, test_cfg of FSF_nuScenes_config.py) = sigmoid, BasePointBBoxCoder.decode on the stage's centres, rotated
        multi-class NMS → (boxes [n,9] (x,y,z,dx,dy,dz,yaw,vx,vy), scores [n], labels [n]).  Needs `refine` to have run."""
        i = self.num_extra_stages - 1
        rois = ops.decode_boxes(st[f"refine{i}_reg"], st[f"refine{i}_centers"])
        boxes, scores, labels, rows = ops.multiclass_nms(rois[:, 1:], st[f"refine{i}_cls"], score_thr, nms_thr, max_num)
        st.update(det_boxes=boxes, det_scores=scores, det_labels=labels, det_rows=rows)
        return boxes, scores, labels

    @torch.no_grad()
    def simple_test(self, points, img_metas, mask_data, mask_anno, score_thr: float = 0.01, nms_thr: float = 0.35,
                    max_num: int = 500, **kwargs):
        """The reference's test entry (FSF.simple_test, FSF.py:1114-1176) with its argument conventions: `points` a list of
        per-sample `[N, 5+3]` tensors (or one tensor), `img_metas` a list of dicts carrying `lidar2img` (list of 4x4),
        `mask_data [B, cams, classes, H, W]`, `mask_anno [B, obj_max_num, 9]`.  Returns mmdet3d's `bbox3d2result` layout, one dict per
        sample with CPU tensors: `boxes_3d [n, 9]` (x, y, z, dx, dy, dz, yaw, vx, vy — the raw tensor a LiDARInstance3DBoxes would
        wrap), `scores_3d [n]`, `labels_3d [n]`.  Samples are processed one after the other (samples_per_gpu = 1 upstream)."""
        if torch.is_tensor(points):
            points = [points]
        results = []
        for b, pts in enumerate(points):
            l2i = img_metas[b]["lidar2img"]
            if not torch.is_tensor(l2i):
                import numpy as np

                l2i = torch.from_numpy(np.asarray([np.asarray(m, dtype=np.float32) for m in l2i], dtype=np.float32))
            l2i = l2i.to(device=pts.device, dtype=torch.float32)
            st = self.forward(pts, mask_data[b], mask_anno[b], l2i)
            st = self.refine(st, pts)
            boxes, scores, labels = self.get_bboxes(st, score_thr, nms_thr, max_num)
            results.append(dict(boxes_3d=boxes.cpu(), scores_3d=scores.cpu(), labels_3d=labels.cpu()))
        return results

    def forward_test(self, points, img_metas, mask_data, mask_anno, **kwargs):
        """FSF.forward_test (FSF.py:1096-1112): the outer lists are test-time augmentations; only the single-view case exists here
        (aug_test is training-repo tooling)."""
        if len(points) != 1:
            raise NotImplementedError("test-time augmentation (aug_test) is outside the hot path")
        return self.simple_test(points[0], img_metas[0], mask_data[0], mask_anno[0], **kwargs)

    # reference checkpoint prefix → attribute here (FSF.__init__ FSF.py:86-164; VoteSegmentor.__init__ single_stage_fsd.py:160-204)
    REFERENCE_PREFIXES = (("segmentor.voxel_encoder.", "voxel_encoder."), ("segmentor.backbone.", "backbone_unet."),
                          ("segmentor.segmentation_head.", "segmentation_head."), ("segmentor.decode_neck.", "decode_neck."))

    def load_reference_state_dict(self, state_dict, strict: bool = False):
        """Load a checkpoint saved by the reference detector: the segmentor's sub-modules live one level down there
        (`segmentor.voxel_encoder.*`, `segmentor.backbone.*` = the sparse U-Net, `segmentor.segmentation_head.*`), every other
        top-level name is the same.  Sparse-convolution kernels stored 5-d are brought to `[27, Cout, Cin]` (spconv 1.x
        `[kz,ky,kx,Cin,Cout]`, spconv 2.x `[Cout,kz,ky,kx,Cin]`; told apart by the shape the target expects).  Returns
        (missing, unexpected) like `load_state_dict`; the internal key names of the un-vendored blocks (SimpleSparseUNet,
        DynamicScatterVFE, SIRLayer) are the recalled ones, so with strict=False check the two lists.  Calls `refresh()`."""
        own = self.state_dict()
        mapped = {}
        odd_kernels = []
        for k, v in state_dict.items():
            for a, b in self.REFERENCE_PREFIXES:
                if k.startswith(a):
                    k = b + k[len(a):]
                    break
            if k in own and v.dim() == 5 and own[k].dim() == 3:
                koff, cout, cin = own[k].shape
                if tuple(v.shape[3:]) == (cin, cout) and v.shape[:3].numel() == koff:        # spconv 1.x
                    v = v.permute(0, 1, 2, 4, 3).reshape(koff, cout, cin)
                elif v.shape[0] == cout and v.shape[4] == cin and v.shape[1:4].numel() == koff:  # spconv 2.x
                    v = v.permute(1, 2, 3, 0, 4).reshape(koff, cout, cin)
                else:
                    odd_kernels.append(f"{k}: {tuple(v.shape)} matches neither spconv layout of [{koff}, {cout}, {cin}]")
            mapped[k] = v
        res = self.load_state_dict(mapped, strict=strict)
        # a non-strict load that leaves parameters at their random initial values must not pass silently
        missing = [k for k in res.missing_keys if own[k].is_floating_point()]
        if missing or res.unexpected_keys or odd_kernels:
            import warnings
            warnings.warn(f"load_reference_state_dict: {len(missing)} floating-point tensor(s) of this model were NOT loaded "
                          f"(first: {missing[:5]}), {len(res.unexpected_keys)} checkpoint key(s) have no counterpart (first: "
                          f"{list(res.unexpected_keys)[:5]}), {len(odd_kernels)} 5-d kernel(s) of unknown layout {odd_kernels[:3]}",
                          RuntimeWarning, stacklevel=2)
        self.refresh()
        return res

    @torch.no_grad()
    def forward(self, points, mask_data, mask_anno, lidar2img):
        stages, st = self.stages(points, mask_data, mask_anno, lidar2img)
        for _, fn in stages:
            fn()
        return st
