"""Host-side mirror of the reference's operator / registry interface for the forward hot path.

Same names, constructor kwargs and forward signatures as the types the stock configs name
(SURVEY.md §8b); all arithmetic goes through the C-ABI ops (fullysparsefusion_b200.ops).  Parameters
live in ordinary nn.Linear / nn.LayerNorm / nn.BatchNorm1d containers laid out exactly as
build_mlp lays them out (`{i}.0.weight`, `{i}.1.*`; sst_ops.py:808-833) so reference checkpoints load;
the tensor-core weight packs are derived lazily and cached (call `.refresh()` after loading weights).
Inference only (eval-mode BatchNorm; no autograd through the kernels).

Types whose source is NOT in the reference tree (DynamicScatterVFE, SIRLayer, DynamicVFELayer,
SimpleSparseUNet — un-vendored mmdet3d fork) follow the published SST/FSD
implementation as recalled; their channel arithmetic is pinned by the configs (5+3+3=11 VFE
inputs, 2x128 cluster features per SIR block → 768, 34 sparse convolutions).
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional

import torch
from torch import nn

from . import ops

# ------------------------------------------------------------------------------------------------
# scatter_v2 + plans
# ------------------------------------------------------------------------------------------------


class ScatterPlan:
    """torch.unique(coors, dim=0) + the CSR every reduction over that ranking reuses
    (scatter_v2's `unq_inv`/`new_coors` pair, sst_ops.py:150-177, plus this framework's rulebook)."""

    def __init__(self, coors: torch.Tensor, lo=None, ext=None, want_index: bool = False):
        res = ops.unique_rows(coors, lo=lo, ext=ext, inv_dtype=torch.int32, return_index=want_index)
        self.new_coors, self.inv32 = res[0], res[1]
        self.index = res[3] if want_index else None
        self.m = self.new_coors.size(0)
        self.csr = ops.build_csr(self.inv32, self.m)
        self.csr.dense = True  # ranks of existing rows: no empty segment
        self._inv64 = None

    @property
    def unq_inv(self) -> torch.Tensor:
        if self._inv64 is None:
            self._inv64 = self.inv32.long()
        return self._inv64

    def reduce(self, feat: torch.Tensor, mode: str) -> torch.Tensor:
        return ops.segment_reduce(feat, self.csr, {"avg": "mean"}.get(mode, mode))

    def gather(self, voxel_feat: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        return ops.gather_rows(voxel_feat, self.inv32, out=out)


def scatter_v2(feat, coors, mode, return_inv=True, min_points=0, unq_inv=None, new_coors=None, plan: ScatterPlan = None):
    """projects/mmdet3d_plugin/ops/sst_ops.py:150-177 (min_points == 0 path)."""
    assert feat.size(0) == coors.size(0)
    assert min_points == 0, "min_points > 0 is not used on the FSF forward path"
    if mode == "avg":
        mode = "mean"
    if plan is None:
        plan = ScatterPlan(coors)
    new_feat = plan.reduce(feat, mode)
    if not return_inv:
        return new_feat, plan.new_coors
    return new_feat, plan.new_coors, plan.unq_inv


# ------------------------------------------------------------------------------------------------
# build_mlp
# ------------------------------------------------------------------------------------------------
class naiveSyncBN1d(nn.BatchNorm1d):
    """Norm type 'naiveSyncBN1d' of the stock configs (FSF_nuScenes_config.py:50,63,85; built through mmcv's build_norm_layer in
    ops/sst_ops.py:814).  Source un-vendored [UPSTREAM-RECALL: detectron2's NaiveSyncBatchNorm for [N, C] inputs]: in training
    with an initialised process group the batch statistics are the rank-mean of (mean, mean of squares) — `dist.naive_sync_bn_stats`,
    one [2C] all-reduce per layer; otherwise exactly nn.BatchNorm1d.  Eval mode (the inference path) is folded into the GEMM
    epilogue by FusedLinear / SparseConvModule and never runs this forward."""

    def forward(self, x):
        import torch.distributed as tdist

        if not self.training or not (tdist.is_available() and tdist.is_initialized() and tdist.get_world_size() > 1):
            return super().forward(x)
        from . import dist as fdist

        mean, var = fdist.naive_sync_bn_stats(x)
        with torch.no_grad():
            m = self.momentum if self.momentum is not None else 0.1
            self.running_mean.mul_(1 - m).add_(mean.detach(), alpha=m)
            self.running_var.mul_(1 - m).add_(var.detach(), alpha=m)
            self.num_batches_tracked += 1
        scale = self.weight * torch.rsqrt(var + self.eps)
        return x * scale + (self.bias - mean * scale)


def build_norm_layer(cfg: dict, num_features: int):
    t = cfg["type"]
    if t == "LN":
        return "ln", nn.LayerNorm(num_features, eps=cfg.get("eps", 1e-5))
    if t == "naiveSyncBN1d":
        return "bn", naiveSyncBN1d(num_features, eps=cfg.get("eps", 1e-5), momentum=cfg.get("momentum", 0.1))
    if t in ("BN1d", "BN", "SyncBN"):
        return "bn", nn.BatchNorm1d(num_features, eps=cfg.get("eps", 1e-5), momentum=cfg.get("momentum", 0.1))
    raise NotImplementedError(t)


def _act_name(act) -> Optional[str]:
    if act is None:
        return None
    act = act.lower()
    if act not in ("relu", "gelu"):
        raise NotImplementedError(f"activation {act!r} is not on the FSF forward path (relu/gelu only)")
    return act


class FusedLinear:
    """One Linear(+norm)(+act) block executed as a single tcgen05 gather-GEMM with fused epilogue."""

    def __init__(self, linear: nn.Linear, norm: Optional[nn.Module], act: Optional[str]):
        self.linear, self.norm, self.act = linear, norm, act
        self._pack = None

    def refresh(self):
        self._pack = None

    def _prepare(self):
        w = self.linear.weight.detach()
        pack = dict(w=ops.gemm_prepack(w.float()), bias=None, norm=None, norm_w=None, norm_b=None, eps=1e-5, wide_ln=False)
        if self.linear.bias is not None:
            pack["bias"] = self.linear.bias.detach().float().contiguous()
        n = self.norm
        if isinstance(n, nn.LayerNorm):
            pack.update(norm="ln", norm_w=n.weight.detach().float().contiguous(), norm_b=n.bias.detach().float().contiguous(),
                        eps=n.eps, wide_ln=w.size(0) > 256)
        elif isinstance(n, nn.BatchNorm1d):  # eval-mode BN (naiveSyncBN1d falls back to it) folded to an affine
            scale = n.weight.detach().float() / torch.sqrt(n.running_var.float() + n.eps)
            shift = n.bias.detach().float() - n.running_mean.float() * scale
            pack.update(norm="affine", norm_w=scale.contiguous(), norm_b=shift.contiguous())
        elif n is not None:
            raise NotImplementedError(type(n))
        self._pack = pack

    def __call__(self, x: torch.Tensor, out: Optional[torch.Tensor] = None, residual=None) -> torch.Tensor:
        if self._pack is None:
            self._prepare()
        p = self._pack
        if p["wide_ln"] and ops.linear_k_splits(x.size(0), self.linear.in_features, self.linear.out_features) > 1 and self.linear.out_features <= 1024:
            # K-split GEMM: its split epilogue sees whole rows, so the wide LayerNorm is fused there (one launch less)
            return ops.gather_gemm(x, p["w"], bias=p["bias"], norm="ln", norm_w=p["norm_w"], norm_b=p["norm_b"], eps=p["eps"],
                                   residual=residual, act=self.act, out=out)
        if p["wide_ln"]:  # LayerNorm wider than one accumulator tile: GEMM, then the row kernel
            y = ops.gather_gemm(x, p["w"])
            return ops.rownorm_act(y, bias=p["bias"], norm="ln", norm_w=p["norm_w"], norm_b=p["norm_b"], eps=p["eps"],
                                   residual=residual, act=self.act, out=out if out is not None else y)
        return ops.gather_gemm(x, p["w"], bias=p["bias"], norm=p["norm"], norm_w=p["norm_w"], norm_b=p["norm_b"],
                               eps=p["eps"], residual=residual, act=self.act, out=out)


class FusedMLP(nn.Sequential):
    """nn.Sequential with build_mlp's module layout; forward runs one fused kernel per block."""

    def __init__(self, *layers, act: Optional[str] = None):
        super().__init__(*layers)
        self._act = act
        self._fused: Optional[List[FusedLinear]] = None

    def refresh(self):
        self._fused = None

    def _build(self):
        fused = []
        for layer in self:
            if isinstance(layer, nn.Linear):
                fused.append(FusedLinear(layer, None, None))
            else:
                mods = list(layer)
                norm = mods[1] if len(mods) > 1 and not isinstance(mods[1], (nn.ReLU, nn.GELU)) else None
                fused.append(FusedLinear(mods[0], norm, self._act))
        self._fused = fused

    def forward(self, x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        if self._fused is None:
            self._build()
        for i, f in enumerate(self._fused):
            x = f(x, out=out if i == len(self._fused) - 1 else None)
        return x


def build_mlp(in_channel, hidden_dims, norm_cfg, is_head=False, act="relu", bias=False, dropout=0) -> FusedMLP:
    """projects/mmdet3d_plugin/ops/sst_ops.py:808-833."""
    assert dropout == 0, "dropout is a no-op at inference and unused by the stock configs"
    act = _act_name(act)
    layers, last = [], in_channel
    for i, c in enumerate(hidden_dims):
        if i == len(hidden_dims) - 1 and is_head:
            layers.append(nn.Linear(last, c, bias=True))
        else:
            layers.append(nn.Sequential(nn.Linear(last, c, bias=bias), build_norm_layer(norm_cfg, c)[1],
                                        nn.ReLU(inplace=True) if act == "relu" else nn.GELU()))
        last = c
    return FusedMLP(*layers, act=act)


# ------------------------------------------------------------------------------------------------
# DynamicScatterVFE
# ------------------------------------------------------------------------------------------------
class DynamicScatterVFE(nn.Module):
    """Registry type 'DynamicScatterVFE' (config FSF_nuScenes_config.py:42-52; call
    single_stage_fsd.py:232): point decoration (xyz - voxel mean, xyz - voxel centre), then
    Linear→BN→ReLU + scatter-max per layer, with the voxel feature concatenated back between layers."""

    def __init__(self, in_channels=4, feat_channels=(), with_distance=False, with_cluster_center=False,
                 with_voxel_center=False, voxel_size=(0.2, 0.2, 4), point_cloud_range=(0, -40, -3, 70.4, 40, 1),
                 norm_cfg=dict(type="BN1d", eps=1e-3, momentum=0.01), mode="max", fusion_layer=None,
                 return_point_feats=False, unique_once=False):
        super().__init__()
        assert not with_distance and fusion_layer is None and mode == "max"
        self.in_channels = in_channels + 3 * int(with_cluster_center) + 3 * int(with_voxel_center)
        self._with_cluster_center, self._with_voxel_center = with_cluster_center, with_voxel_center
        self.return_point_feats, self.unique_once = return_point_feats, unique_once
        self.voxel_size, self.point_cloud_range = list(voxel_size), list(point_cloud_range)
        chans = [self.in_channels] + list(feat_channels)
        layers = []
        for i in range(len(chans) - 1):
            cin = chans[i] * (2 if i > 0 else 1)
            layers.append(nn.Sequential(nn.Linear(cin, chans[i + 1], bias=False), build_norm_layer(norm_cfg, chans[i + 1])[1],
                                        nn.ReLU(inplace=True)))
        self.vfe_layers = nn.ModuleList(layers)
        self._fused = None

    def refresh(self):
        self._fused = None

    def forward(self, features, coors, points=None, img_feats=None, img_metas=None, return_inv=False,
                plan: Optional[ScatterPlan] = None):
        if self._fused is None:
            self._fused = [FusedLinear(l[0], l[1], "relu") for l in self.vfe_layers]
        if plan is None:
            plan = ScatterPlan(coors)
        n, cin = features.shape
        dec = torch.empty((n, self.in_channels), dtype=torch.float32, device=features.device)
        voxel_mean = plan.reduce(features, "mean") if self._with_cluster_center else None
        ops.vfe_decorate(features, coors, plan.inv32, voxel_mean, self.voxel_size, self.point_cloud_range,
                         self._with_cluster_center, self._with_voxel_center, dec)
        x = dec
        voxel_feats = None
        for i, f in enumerate(self._fused):
            c = f.linear.out_features
            last = i == len(self._fused) - 1
            buf = torch.empty((n, c if last else 2 * c), dtype=torch.float32, device=x.device)
            point_feats = f(x, out=buf[:, :c])
            voxel_feats = plan.reduce(point_feats, "max")
            if not last:
                plan.gather(voxel_feats, out=buf[:, c:])   # torch.cat([point_feats, feat_per_point], 1)
                x = buf
        if self.return_point_feats:
            return point_feats
        if return_inv:
            return voxel_feats, plan.new_coors, plan.unq_inv
        return voxel_feats, plan.new_coors


# ------------------------------------------------------------------------------------------------
# SIR
# ------------------------------------------------------------------------------------------------
class DynamicVFELayer(nn.Module):
    """Linear(bias=False) → norm → act (the per-point layer of SIRLayer)."""

    def __init__(self, in_channels, out_channels, norm_cfg=dict(type="BN1d", eps=1e-3, momentum=0.01), act="relu", dropout=0.0):
        super().__init__()
        self.norm = build_norm_layer(norm_cfg, out_channels)[1]
        self.linear = nn.Linear(in_channels, out_channels, bias=False)
        self.act = _act_name(act)
        self._fused = None

    def refresh(self):
        self._fused = None

    def forward(self, x, out=None):
        if self._fused is None:
            self._fused = FusedLinear(self.linear, self.norm, self.act)
        return self._fused(x, out=out)


class SIRLayer(nn.Module):
    """Registry type 'SIRLayer' (built by SIR, models/backbones/sir.py:41-62; called :78,81).

    forward(features [n, 3+C], coors [n, D], f_cluster [n, 3]) with rel_mlp gating:
      x = cat(xyz / xyz_normalizer, feats) * rel_mlp(f_cluster / rel_dist_scaler)
      for each DynamicVFELayer: point = layer(x); cluster = scatter_max(point); x = cat(point, cluster[inv])
      cluster_feats = cat(all layers' cluster maxima); residual on the point features when shapes match."""

    def __init__(self, in_channels=4, feat_channels=(), with_distance=False, with_cluster_center=False, with_rel_mlp=True,
                 rel_feat_dim=16, rel_mlp_hidden_dims=(16,), rel_mlp_in_channel=3, with_voxel_center=False,
                 voxel_size=(0.2, 0.2, 4), point_cloud_range=(0, -40, -3, 70.4, 40, 1),
                 norm_cfg=dict(type="BN1d", eps=1e-3, momentum=0.01), mode="max", fusion_layer=None, return_point_feats=False,
                 return_inv=False, rel_dist_scaler=1.0, with_shortcut=True, xyz_normalizer=(1.0, 1.0, 1.0), act="relu", dropout=0.0):
        super().__init__()
        assert not with_distance and not with_cluster_center and not with_voxel_center and mode == "max"
        self.in_channels = in_channels
        self.with_rel_mlp, self.return_point_feats = with_rel_mlp, return_point_feats
        self.rel_dist_scaler, self.with_shortcut = float(rel_dist_scaler), with_shortcut
        self.xyz_normalizer = [float(v) for v in xyz_normalizer]
        if with_rel_mlp:
            self.rel_mlp = build_mlp(rel_mlp_in_channel, list(rel_mlp_hidden_dims) + [in_channels], norm_cfg, act=act)
        chans = [in_channels] + list(feat_channels)
        self.vfe_layers = nn.ModuleList(
            [DynamicVFELayer(chans[i] * (2 if i > 0 else 1), chans[i + 1], norm_cfg, act=act) for i in range(len(chans) - 1)])
        self.num_vfe = len(self.vfe_layers)

    def refresh(self):
        self._gate_pack = None

    def _fused_gate(self):
        """(layers, eps, act) for ops.sir_gate_input when rel_mlp is the stock 3-block LN stack that kernel covers."""
        if not self.with_rel_mlp:
            return None
        pack = getattr(self, "_gate_pack", None)
        if pack is None:
            blocks = list(self.rel_mlp)
            ok = (len(blocks) == 3 and all(isinstance(b, nn.Sequential) and isinstance(b[1], nn.LayerNorm) and b[0].bias is None
                                           for b in blocks)
                  and blocks[0][0].in_features == 3 and blocks[0][0].out_features <= 32 and blocks[1][0].out_features <= 32
                  and self.in_channels <= 256 and len({b[1].eps for b in blocks}) == 1)
            if ok:
                layers = [(b[0].weight.detach().float().contiguous(), b[1].weight.detach().float().contiguous(),
                           b[1].bias.detach().float().contiguous()) for b in blocks]
                pack = (layers, blocks[0][1].eps, self.rel_mlp._act)
            else:
                pack = False
            self._gate_pack = pack
        return pack or None

    def forward(self, features, coors, f_cluster=None, points=None, img_feats=None, img_metas=None, return_both=False,
                unq_inv_once=None, new_coors_once=None, plan: Optional[ScatterPlan] = None, features_b=None):
        """features_b (extension): the input rows are cat(features, features_b) — lets SIR skip its torch.cat."""
        if plan is None:
            plan = ScatterPlan(coors)
        n = features.size(0)
        dev = features.device
        fused = self._fused_gate()
        if fused is not None:   # 3 → h1 → h2 → Cin gate MLP + normalisation + multiply in one kernel
            x = ops.sir_gate_input(features, f_cluster, self.rel_dist_scaler, self.xyz_normalizer, fused[0], fused[1], fused[2],
                                   features_b=features_b)
        else:
            if features_b is not None:
                features = torch.cat([features, features_b], 1)
            gate = self.rel_mlp(ops.div_cols(f_cluster, [self.rel_dist_scaler] * f_cluster.size(1))) if self.with_rel_mlp else None
            x = ops.sir_input(features, self.xyz_normalizer, gate)   # cat(xyz/norm, feats) * gate, one pass
        ori = x
        cluster_list = []
        point_feats = None
        for i, layer in enumerate(self.vfe_layers):
            c = layer.linear.out_features
            last = i == self.num_vfe - 1
            buf = torch.empty((n, c if last else 2 * c), dtype=torch.float32, device=dev)
            point_feats = layer(x, out=buf[:, :c])
            cluster = plan.reduce(point_feats, "max")
            cluster_list.append(cluster)
            if not last:
                plan.gather(cluster, out=buf[:, c:])
                x = buf
        cluster_feats = torch.cat(cluster_list, dim=1) if len(cluster_list) > 1 else cluster_list[0]
        if return_both or self.return_point_feats:
            if self.with_shortcut and point_feats.shape == ori.shape:
                point_feats = ops.add_(point_feats, ori)
            if return_both:
                return point_feats, cluster_feats, plan.new_coors
            return point_feats, cluster_feats
        return cluster_feats, plan.new_coors


class DynamicClusterVFE(SIRLayer):
    """Registry type 'DynamicClusterVFE' (VOXEL_ENCODERS): the block the reference's FullySparseBboxHead builds through
    builder.build_voxel_encoder (models/roi_heads/bbox_heads/fsd_bbox_head.py:62-87) and calls as
    block(in_feats, roi_inds [P], f_cluster, unq_inv_once=, new_coors_once=) (:135,140).  Same computation as SIRLayer; it takes
    the extra constructor kwargs of that call site (fusion='cat', pos_fusion='mul', cat_voxel_feats=True: the only combination
    the configs use), accepts 1-d group ids and returns the group coordinates in the shape / dtype they came in (the head indexes
    RoI slots with them, :178-197).  The once-computed unique of the caller is not needed: the ranking is recomputed here (or
    shared through `plan`)."""

    def __init__(self, *args, fusion="cat", pos_fusion="mul", cat_voxel_feats=True, **kwargs):
        if fusion != "cat" or pos_fusion != "mul" or not cat_voxel_feats:
            raise NotImplementedError("DynamicClusterVFE: only fusion='cat', pos_fusion='mul', cat_voxel_feats=True is on the FSF path")
        super().__init__(*args, **kwargs)

    def forward(self, features, coors, f_cluster=None, points=None, img_feats=None, img_metas=None, return_both=False,
                unq_inv_once=None, new_coors_once=None, plan: Optional[ScatterPlan] = None, features_b=None):
        one_d = coors.dim() == 1
        if plan is None:
            plan = ScatterPlan(coors.view(-1, 1) if one_d else coors)
        out = super().forward(features, coors, f_cluster, return_both=return_both, plan=plan, features_b=features_b)
        if one_d and (return_both or not self.return_point_feats):   # the last element is the group coordinates
            out = out[:-1] + (out[-1].view(-1).to(coors.dtype),)
        return out


class SIR(nn.Module):
    """models/backbones/sir.py:14-85 (same constructor and forward)."""

    def __init__(self, num_blocks=5, in_channels=(), feat_channels=(), rel_mlp_hidden_dims=(), with_rel_mlp=True,
                 with_distance=False, with_cluster_center=False, norm_cfg=dict(type="LN", eps=1e-3), mode="max",
                 xyz_normalizer=(1.0, 1.0, 1.0), act="relu", dropout=0, unique_once=False):
        super().__init__()
        self.num_blocks, self.unique_once = num_blocks, unique_once
        self.block_list = nn.ModuleList([
            SIRLayer(in_channels=in_channels[i], feat_channels=feat_channels[i], with_distance=with_distance,
                     with_cluster_center=with_cluster_center, with_rel_mlp=with_rel_mlp,
                     rel_mlp_hidden_dims=rel_mlp_hidden_dims[i], with_voxel_center=False, norm_cfg=norm_cfg, mode=mode,
                     return_point_feats=i != num_blocks - 1, rel_dist_scaler=10.0, xyz_normalizer=xyz_normalizer, act=act,
                     dropout=dropout) for i in range(num_blocks)])

    def forward(self, points, features, coors, f_cluster=None, plan: Optional[ScatterPlan] = None):
        if plan is None:
            plan = ScatterPlan(coors)        # unique_once: one ranking shared by all blocks (sir.py:67-70)
        out_feats = features
        cluster_feat_list = []
        out_coors = None
        for i, block in enumerate(self.block_list):
            # in_feats = torch.cat([points, out_feats], 1) (sir.py:76), read from its two sources by the gate kernel
            if i < self.num_blocks - 1:
                out_feats, c = block(points, coors, f_cluster, plan=plan, features_b=out_feats)
            else:
                out_feats, c, out_coors = block(points, coors, f_cluster, return_both=True, plan=plan, features_b=out_feats)
            cluster_feat_list.append(c)
        return out_feats, torch.cat(cluster_feat_list, dim=1), out_coors


# ------------------------------------------------------------------------------------------------
# SimpleSparseUNet
# ------------------------------------------------------------------------------------------------
class Rulebook:
    """Neighbour table of one indice_key plus the mask-sorted row order the gather-GEMM tiles walk."""
    __slots__ = ("nbr", "order", "nbr_ro")

    def __init__(self, nbr: torch.Tensor, sort_rows: bool = True):
        self.nbr = nbr
        # rows with the same set of present offsets become adjacent: a 128-row tile then skips every offset none of
        # its rows has (21 → ~11 of 27 offsets per tile on LiDAR voxel sets); results do not depend on the order
        # (the sort is ~12 small launches: only worth it for the big levels)
        self.order = ops.rulebook_row_order(nbr) if sort_rows and nbr.size(0) > 1 and nbr.size(1) >= 50000 else None
        # the table once more in that order (tile tables become contiguous runs for the kernel's scheduler warp)
        self.nbr_ro = ops.permute_rulebook(nbr, self.order) if self.order is not None else None


class SparseConvModule(nn.Module):
    """conv (bias=False) → BN1d → ReLU, the ('conv','norm','act') block of make_sparse_convmodule.
    weight: [27, cout, cin], offsets ordered (kz,ky,kx) — see INTEGRATION.md for the spconv layouts."""

    def __init__(self, cin, cout, norm_cfg, conv_type="SubMConv3d", stride=1, padding=1, indice_key=None, act=True):
        super().__init__()
        self.conv_type, self.stride, self.padding, self.indice_key = conv_type, stride, padding, indice_key
        self.weight = nn.Parameter(torch.empty(27, cout, cin))
        nn.init.kaiming_uniform_(self.weight.view(27 * cout, cin), a=5 ** 0.5)
        with torch.no_grad():
            self.weight.mul_(27 ** -0.5)
        self.bn = build_norm_layer(norm_cfg, cout)[1]
        self.act = "relu" if act else None
        self._pack = None

    def refresh(self):
        self._pack = None

    def forward(self, feats, rb, out=None, residual=None, residual_post=False):
        nbr, order, nbr_ro = (rb.nbr, rb.order, rb.nbr_ro) if isinstance(rb, Rulebook) else (rb, None, None)
        if self._pack is None:
            n = self.bn
            scale = n.weight.detach().float() / torch.sqrt(n.running_var.float() + n.eps)
            shift = n.bias.detach().float() - n.running_mean.float() * scale
            self._pack = (ops.gemm_prepack(self.weight.detach().float()), scale.contiguous(), shift.contiguous())
        w, scale, shift = self._pack
        return ops.gather_gemm(feats, w, nbr=nbr, norm="affine", norm_w=scale, norm_b=shift, residual=residual,
                               act=self.act, out=out, residual_post=residual_post, row_order=order, nbr_ro=nbr_ro)


class SparseBasicBlock(nn.Module):
    """conv-bn-relu, conv-bn, += identity, relu (mmdet3d SparseBasicBlock)."""

    def __init__(self, c, norm_cfg, indice_key):
        super().__init__()
        self.conv1 = SparseConvModule(c, c, norm_cfg, indice_key=indice_key)
        self.conv2 = SparseConvModule(c, c, norm_cfg, indice_key=indice_key)

    def forward(self, x, nbr, out=None):
        return self.conv2(self.conv1(x, nbr), nbr, out=out, residual=x)


class SimpleSparseUNet(nn.Module):
    """Registry type 'SimpleSparseUNet' (config FSF_nuScenes_config.py:58-70; call single_stage_fsd.py:234).
    forward(voxel_info) -> [{'voxel_feats': [M, C]}]; voxel rows must be in ranked (lexicographic) order."""

    def __init__(self, in_channels, sparse_shape, order=("conv", "norm", "act"),
                 norm_cfg=dict(type="BN1d", eps=1e-3, momentum=0.01), base_channels=16, output_channels=128, ndim=3,
                 encoder_channels=((16,), (32, 32, 32), (64, 64, 64), (64, 64, 64)),
                 encoder_paddings=((1,), (1, 1, 1), (1, 1, 1), ((0, 1, 1), 1, 1)),
                 decoder_channels=((64, 64, 64), (64, 64, 32), (32, 32, 16), (16, 16, 16)),
                 decoder_paddings=((1, 0), (1, 0), (0, 0), (0, 1)), keep_coors_dims=None, act_type="relu", init_cfg=None):
        super().__init__()
        assert tuple(order) == ("conv", "norm", "act") and ndim == 3 and act_type == "relu"
        self.sparse_shape = [int(v) for v in sparse_shape]
        self.stage_num = len(encoder_channels)
        self.encoder_channels, self.encoder_paddings = encoder_channels, encoder_paddings
        self.conv_input = SparseConvModule(in_channels, base_channels, norm_cfg, indice_key="subm1")
        cin = base_channels
        self.encoder_layers = nn.ModuleList()
        for i, blocks in enumerate(encoder_channels):
            stage = nn.ModuleList()
            for j, cout in enumerate(blocks):
                pad = tuple(encoder_paddings[i])[j]
                if i != 0 and j == 0:
                    stage.append(SparseConvModule(cin, cout, norm_cfg, "SparseConv3d", 2, pad, f"spconv{i + 1}"))
                else:
                    stage.append(SparseConvModule(cin, cout, norm_cfg, "SubMConv3d", 1, pad, f"subm{i + 1}"))
                cin = cout
            self.encoder_layers.append(stage)
        n = len(decoder_channels)
        for i, bc in enumerate(decoder_channels):
            lvl = n - i
            setattr(self, f"lateral_layer{lvl}", SparseBasicBlock(cin, norm_cfg, f"subm{lvl}"))
            assert bc[0] == cin
            setattr(self, f"merge_layer{lvl}", SparseConvModule(cin * 2, bc[1], norm_cfg, indice_key=f"subm{lvl}"))
            if lvl != 1:
                setattr(self, f"upsample_layer{lvl}", SparseConvModule(cin, bc[2], norm_cfg, "SparseInverseConv3d", 2,
                                                                       None, f"spconv{lvl}"))
            else:
                setattr(self, f"upsample_layer{lvl}", SparseConvModule(cin, bc[2], norm_cfg, indice_key="subm1"))
            assert bc[1] == cin
            cin = bc[2]
        self.output_channels = cin

    @staticmethod
    def _triple(p):
        return [int(p)] * 3 if isinstance(p, int) else [int(v) for v in p]

    def build_rulebooks(self, coors32: torch.Tensor, index: "ops.VoxelIndex", batch_size: int):
        """All neighbour tables of the U-Net, one per indice_key (spconv reuses them by key)."""
        shape = [batch_size] + self.sparse_shape
        levels = [dict(coors=coors32, index=index, shape=shape)]
        rb: Dict[str, Rulebook] = {"subm1": Rulebook(ops.conv_rulebook(coors32, index, 3, 1, 1))}
        for i in range(1, self.stage_num):
            pad = self._triple(tuple(self.encoder_paddings[i])[0])
            prev = levels[-1]
            oshape = [batch_size] + [(prev["shape"][1 + a] + 2 * pad[a] - 3) // 2 + 1 for a in range(3)]
            oc, oindex = ops.conv_out_index(prev["coors"], oshape, 3, 2, pad)
            rb[f"spconv{i + 1}"] = Rulebook(ops.conv_rulebook(oc, prev["index"], 3, 2, pad))
            rb[f"spconv{i + 1}_inv"] = Rulebook(ops.conv_rulebook(prev["coors"], oindex, 3, 2, pad, transposed=True))
            rb[f"subm{i + 1}"] = Rulebook(ops.conv_rulebook(oc, oindex, 3, 1, 1))
            levels.append(dict(coors=oc, index=oindex, shape=oshape))
        return rb, levels

    def forward(self, voxel_info, rulebooks=None):
        feats = voxel_info["voxel_feats"]
        coors = voxel_info["voxel_coors"]
        if rulebooks is None:
            coors32 = coors.int().contiguous()
            batch_size = int(voxel_info.get("batch_size", 0)) or int(coors32[:, 0].max().item()) + 1
            index = voxel_info.get("voxel_index")
            if index is None:
                shape = [batch_size] + self.sparse_shape
                index = ops.unique_rows(coors32, lo=[0, 0, 0, 0], ext=shape, return_unique=False, return_index=True)[3]
            rulebooks, _ = self.build_rulebooks(coors32, index, batch_size)
        rb = rulebooks
        x = self.conv_input(feats, rb["subm1"])
        enc = []
        for i, stage in enumerate(self.encoder_layers):
            for layer in stage:
                x = layer(x, rb[layer.indice_key])
            enc.append(x)
        # decoder: every level's [x_bottom | lateral] concatenation is written in place (the upsample conv of the
        # level above stores straight into the left half), and x_merge + reduce_channel(x) is the merge conv's epilogue
        cat = None
        for lvl in range(self.stage_num, 0, -1):
            lat_in = enc[lvl - 1]
            n, c = lat_in.shape
            if cat is None:   # top level: x_bottom is the last encoder output itself
                cat = torch.empty((n, 2 * c), dtype=torch.float32, device=lat_in.device)
                cat[:, :c].copy_(enc[-1])
            getattr(self, f"lateral_layer{lvl}")(lat_in, rb[f"subm{lvl}"], out=cat[:, c:])
            merge = getattr(self, f"merge_layer{lvl}")
            reduced = ops.reduce_channel(cat, merge.weight.size(1))   # reduce_channel(x, C): sum of channel groups
            x = merge(cat, rb[f"subm{lvl}"], residual=reduced, residual_post=True)   # x_merge.features + x.features
            up = getattr(self, f"upsample_layer{lvl}")
            if lvl != 1:
                n_next, c_next = enc[lvl - 2].shape
                assert up.weight.size(1) == c_next
                cat = torch.empty((n_next, 2 * c_next), dtype=torch.float32, device=x.device)
                up(x, rb[f"spconv{lvl}_inv"], out=cat[:, :c_next])
            else:
                x = up(x, rb["subm1"])
        return [{"voxel_feats": x}]


# ------------------------------------------------------------------------------------------------
# neck, segmentation head
# ------------------------------------------------------------------------------------------------
class Voxel2PointScatterNeck(nn.Module):
    """models/necks/voxel2point_neck.py:14-70."""

    def __init__(self, point_cloud_range=None, voxel_size=None, with_xyz=True, normalize_local_xyz=False):
        super().__init__()
        assert with_xyz and not normalize_local_xyz
        self.point_cloud_range, self.voxel_size = list(point_cloud_range), list(voxel_size)

    def forward(self, points, pts_coors, voxel_feats, voxel2point_inds, voxel_padding=-1):
        assert points.size(0) == pts_coors.size(0) == voxel2point_inds.size(-1)
        out, mask = ops.voxel2point_neck(points, pts_coors, voxel_feats, voxel2point_inds, self.voxel_size,
                                         self.point_cloud_range, float(voxel_padding))
        return out, mask


class VoteSegHead(nn.Module):
    """models/decode_heads/segmentation_head.py:22-104 (forward + decode_vote_targets only)."""

    def __init__(self, in_channel, num_classes, hidden_dims=(), dropout_ratio=0.0, conv_cfg=None,
                 norm_cfg=dict(type="naiveSyncBN1d"), act_cfg=dict(type="ReLU"), **unused):
        super().__init__()
        self.num_classes = num_classes + 1  # background (segmentation_head.py:58-60)
        act = act_cfg["type"].lower()
        self.pre_seg_conv = build_mlp(in_channel, list(hidden_dims), norm_cfg, act=act) if len(hidden_dims) else None
        end = hidden_dims[-1] if len(hidden_dims) else in_channel
        self.conv_seg = nn.Linear(end, self.num_classes)
        self.voting = nn.Linear(end, self.num_classes * 3)
        self._heads = None

    def refresh(self):
        self._heads = None
        if self.pre_seg_conv is not None:
            self.pre_seg_conv.refresh()

    def forward(self, voxel_feat):
        if self._heads is None:
            self._heads = (FusedLinear(self.conv_seg, None, None), FusedLinear(self.voting, None, None))
        x = self.pre_seg_conv(voxel_feat) if self.pre_seg_conv is not None else voxel_feat
        return self._heads[0](x), self._heads[1](x)

    @staticmethod
    def decode_vote_targets(preds):
        return ops.vote_decode(preds)


# ------------------------------------------------------------------------------------------------
# DynamicPointROIExtractor (query refinement)
# ------------------------------------------------------------------------------------------------
class DynamicPointROIExtractor(nn.Module):
    """models/roi_heads/roi_extractors/dynamic_point_roi_extractor.py:11-100 (config FSF_nuScenes_config.py:290-294).
    forward(pts_xyz [N,3+], batch_inds [N], rois [K,8] = (batch, x,y,z,w,l,h,rz)) ->
        (point ids [P] i64, roi ids [P] i64, dict(local_xyz [P,3], boundary_offset [P,6], is_in_margin [P]))
    One sample per call here (samples_per_gpu = 1); an empty result keeps one fake row of -1 ids as upstream
    (dynamic_point_pool_op.py:36-40)."""

    def __init__(self, init_cfg=None, debug=True, extra_wlh=(0, 0, 0), max_inbox_point=512, max_all_pts=50000):
        super().__init__()
        self.debug, self.extra_wlh, self.max_inbox_point, self.max_all_pts = debug, list(extra_wlh), max_inbox_point, max_all_pts

    @torch.no_grad()
    def forward(self, pts_xyz: torch.Tensor, batch_inds: torch.Tensor, rois: torch.Tensor):
        assert len(pts_xyz) > 0 and len(rois) > 0 and (batch_inds is None or len(batch_inds) > 0)   # one sample per call
        dev = pts_xyz.device
        cap = self.max_all_pts
        out_pts_idx = torch.full((cap,), -1, dtype=torch.int64, device=dev)
        out_roi_idx = torch.full((cap,), -1, dtype=torch.int64, device=dev)
        out_feats = torch.zeros((cap, 13), dtype=torch.float32, device=dev)
        num = ops.dynamic_point_pool(rois[:, 1:8], pts_xyz, self.extra_wlh, self.max_inbox_point, out_pts_idx, out_roi_idx,
                                     out_feats)
        p = max(int(num.item()), 1)   # the output-size read the reference performs as a boolean mask (:34-44)
        inds, roi_inds, info = out_pts_idx[:p], out_roi_idx[:p], out_feats[:p]
        return inds, roi_inds, dict(local_xyz=info[:, 3:6], boundary_offset=info[:, 6:12], is_in_margin=info[:, 12])


# ------------------------------------------------------------------------------------------------
# FullySparseBboxHead (query refinement)
# ------------------------------------------------------------------------------------------------
class FullySparseBboxHead(nn.Module):
    """models/roi_heads/bbox_heads/fsd_bbox_head.py:22-151 (config single_refine_sir_layer, FSF_nuScenes_config.py:296-320):
    num_blocks DynamicClusterVFE blocks (= SIRLayer with a 13-input relative-position MLP) over the RoI ids of the pooled
    points; returns (roi feats [num_rois, sum of block widths], nonempty_roi_mask [num_rois]).

    forward(pts_xyz [P,5], pts_features [P,C], pts_info dict(local_xyz, boundary_offset, is_in_margin), roi_inds [P] i64/i32,
            rois [K,8] = (batch, x,y,z,w,l,h,rz)).  A roi id of -1 (the extractor's fake row) forms group -1, dropped by the
    alignment step exactly as upstream (:178-197)."""

    def __init__(self, num_classes=10, num_blocks=3, in_channels=(), feat_channels=(), with_distance=False, with_cluster_center=False,
                 with_rel_mlp=True, rel_mlp_hidden_dims=(), rel_mlp_in_channels=(), reg_mlp=None, cls_mlp=None, mode="max",
                 xyz_normalizer=(20, 20, 4), cat_voxel_feats=True, pos_fusion="mul", fusion="cat", act="gelu", geo_input=True,
                 use_middle_cluster_feature=True, norm_cfg=dict(type="LN", eps=1e-3, momentum=0.01), dropout=0, unique_once=False,
                 init_cfg=None, no_head=None):
        super().__init__()
        assert mode == "max" and pos_fusion == "mul" and fusion == "cat" and cat_voxel_feats and not dropout
        self.num_blocks, self.geo_input, self.use_middle_cluster_feature = num_blocks, geo_input, use_middle_cluster_feature
        self.block_list = nn.ModuleList([
            SIRLayer(in_channels=in_channels[i], feat_channels=feat_channels[i], with_rel_mlp=with_rel_mlp,
                     rel_mlp_hidden_dims=rel_mlp_hidden_dims[i], rel_mlp_in_channel=rel_mlp_in_channels[i], norm_cfg=norm_cfg, mode=mode,
                     return_point_feats=i != num_blocks - 1, rel_dist_scaler=10.0, xyz_normalizer=xyz_normalizer, act=act)
            for i in range(num_blocks)])

    @torch.no_grad()
    def forward(self, pts_xyz, pts_features, pts_info, roi_inds, rois):
        assert pts_features.size(0) > 0
        dev = pts_xyz.device
        num_rois = rois.size(0)
        ids32 = roi_inds.to(torch.int32).contiguous()
        # rel_xyz = pts_xyz[:, :3] - roi_centers[roi_inds]  (:112): the fake row (-1) reads the last roi, as torch indexing does
        idx = torch.where(ids32 < 0, ids32 + num_rois, ids32)
        rel_xyz, _ = ops.cluster_delta(pts_xyz, rois[:, 1:4].contiguous(), idx)
        f_cluster = torch.cat([pts_info["local_xyz"], pts_info["boundary_offset"], pts_info["is_in_margin"][:, None], rel_xyz], dim=-1)
        plan = ScatterPlan(ids32.view(-1, 1))                      # unique_once: one ranking of the roi ids for all blocks
        geo = ops.div_cols(f_cluster, [10.0] * f_cluster.size(1)) if self.geo_input else None
        out_feats = pts_features
        cluster_feat_list = []
        out_coors = None
        for i, block in enumerate(self.block_list):
            parts = [pts_xyz, out_feats] + ([geo] if geo is not None else [])
            in_feats = torch.cat(parts, 1)
            if i < self.num_blocks - 1:
                out_feats, c = block(in_feats, ids32, f_cluster, plan=plan)
                if self.use_middle_cluster_feature:
                    cluster_feat_list.append(c)
            else:
                c, out_coors = block(in_feats, ids32, f_cluster, plan=plan)
                cluster_feat_list.append(c)
        feats = torch.cat(cluster_feat_list, dim=1)
        return self.align(feats, out_coors.view(-1), num_rois)

    @staticmethod
    def align(feats: torch.Tensor, coors: torch.Tensor, num_rois: int):
        """get_nonempty_roi_mask + align_roi_feature_and_rois (fsd_bbox_head.py:152-197): group rows → their RoI slots; the
        group of the fake row (-1) is dropped, RoIs without points stay zero."""
        dev = feats.device
        keep = ops.compact_indices(coors >= 0)
        new_feature = torch.zeros((num_rois, feats.size(1)), dtype=torch.float32, device=dev)
        mask = torch.zeros(num_rois, dtype=torch.bool, device=dev)
        if keep.numel():
            dst = ops.gather_int_rows(coors.to(torch.int32).view(-1, 1), keep).view(-1).long()
            new_feature[dst] = ops.gather_rows(feats, keep)
            mask[dst] = True
        return new_feature, mask
