"""Training step of the segmentation stage (BASELINE configs[2] "DDP train on 8 GPUs").

What trains here is the first stage of FSF — `VoteSegmentor` of the reference (models/detectors/single_stage_fsd.py:206-240:
DynamicScatterVFE → SimpleSparseUNet → Voxel2PointScatterNeck → VoteSegHead) — with the SAME parameters the inference path
uses (`FSF.voxel_encoder`, `.backbone_unet`, `.decode_neck`, `.segmentation_head`), so a state dict trained here loads into the
detector unchanged.  Training runs the unfused forms: every Linear / sparse convolution is `autograd.sparse_conv` (forward and
input gradient on the tcgen05 gather-GEMM, weight gradient on `fsfb_conv_wgrad`), BatchNorm layers are the modules themselves in
training mode (`naiveSyncBN1d`: one differentiable [2C] all-reduce per layer), scatter-max is the `torch_scatter` shim with its
argmax-routed gradient.  Losses follow the reference head (models/decode_heads/segmentation_head.py:106-240): sigmoid focal loss
on the class logits (FocalLoss gamma 2, alpha 0.25, FSF_nuScenes_config.py:96-102) and an L1 loss on the votes of foreground
points.  Labels are synthetic (bench) or the caller's.

`BucketedReducer` is the data-parallel half: gradients are copied into flat fp32 buckets in reverse registration order as autograd
produces them (post-accumulate hooks); a full bucket starts ONE asynchronous NCCL all-reduce on a side stream while backward
continues (tools/train.py:244-251 wraps the model in MMDistributedDataParallel, which does the same with 25 MB buckets); `finish()`
waits, averages and scatters the views back.  The share of the step spent waiting for the last bucket is what `bench.py --train`
reports as `allreduce_exposed_ms`.

Not covered: the detection heads' losses and target assignment (FSF.forward_train :596-655), the SIR / refine stages' backward."""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

from . import autograd as AG
from . import modules as M
from . import ops
from .shims import torch_scatter as TS


# ------------------------------------------------------------------------------------------------
# unfused, differentiable forwards of the inference modules (same parameters)
# ------------------------------------------------------------------------------------------------
def linear_train(lin: nn.Linear, x: torch.Tensor) -> torch.Tensor:
    y = AG.sparse_conv(x, lin.weight, None)
    return y if lin.bias is None else y + lin.bias


def mlp_train(mlp: nn.Sequential, x: torch.Tensor) -> torch.Tensor:
    """build_mlp's layout (ops/sst_ops.py:808-833): Sequential(Linear, norm, act) blocks and, for heads, a last plain Linear."""
    for layer in mlp:
        if isinstance(layer, nn.Linear):
            x = linear_train(layer, x)
        else:
            for mod in layer:
                x = linear_train(mod, x) if isinstance(mod, nn.Linear) else mod(x)
    return x


def vfe_train(vfe: M.DynamicScatterVFE, features: torch.Tensor, coors: torch.Tensor, plan: M.ScatterPlan) -> torch.Tensor:
    """DynamicScatterVFE.forward (config FSF_nuScenes_config.py:42-52): the decoration is a constant of the input; the layers are
    Linear → BN → ReLU + scatter-max, the voxel maximum concatenated back to the points between layers."""
    n = features.size(0)
    dec = torch.empty((n, vfe.in_channels), dtype=torch.float32, device=features.device)
    with torch.no_grad():
        voxel_mean = plan.reduce(features, "mean") if vfe._with_cluster_center else None
        ops.vfe_decorate(features, coors, plan.inv32, voxel_mean, vfe.voxel_size, vfe.point_cloud_range, vfe._with_cluster_center,
                         vfe._with_voxel_center, dec)
    inv = plan.unq_inv
    x, voxel = dec, None
    for i, layer in enumerate(vfe.vfe_layers):
        p = F.relu(layer[1](linear_train(layer[0], x)))
        voxel, _ = TS.scatter_max(p, inv, dim=0, dim_size=plan.m)
        if i != len(vfe.vfe_layers) - 1:
            x = torch.cat([p, voxel[inv]], dim=1)
    return voxel


def conv_train(m: M.SparseConvModule, x: torch.Tensor, rb, residual: Optional[torch.Tensor] = None) -> torch.Tensor:
    """conv → BN → (+ residual) → ReLU, the block `SparseConvModule` folds into one epilogue at inference."""
    if isinstance(rb, M.Rulebook):
        y = AG.sparse_conv(x, m.weight, rb.nbr, rb.order, rb.nbr_ro, symmetric=m.conv_type == "SubMConv3d")
    else:
        y = AG.sparse_conv(x, m.weight, rb)
    y = m.bn(y)
    if residual is not None:
        y = y + residual
    return F.relu(y) if m.act else y


def unet_train(net: M.SimpleSparseUNet, feats: torch.Tensor, rb: Dict[str, "M.Rulebook"]) -> torch.Tensor:
    """SimpleSparseUNet.forward, written with torch.cat instead of the in-place concatenation buffers of the inference path."""
    x = conv_train(net.conv_input, feats, rb["subm1"])
    enc = []
    for stage in net.encoder_layers:
        for layer in stage:
            x = conv_train(layer, x, rb[layer.indice_key])
        enc.append(x)
    bottom = enc[-1]
    for lvl in range(net.stage_num, 0, -1):
        lat_in = enc[lvl - 1]
        block = getattr(net, f"lateral_layer{lvl}")
        lat = conv_train(block.conv2, conv_train(block.conv1, lat_in, rb[f"subm{lvl}"]), rb[f"subm{lvl}"], residual=lat_in)
        cat = torch.cat([bottom, lat], dim=1)
        merge = getattr(net, f"merge_layer{lvl}")
        c_out = merge.weight.size(1)
        reduced = cat.view(cat.size(0), c_out, -1).sum(dim=2)          # reduce_channel: sums of consecutive channel groups
        x = conv_train(merge, cat, rb[f"subm{lvl}"]) + reduced
        up = getattr(net, f"upsample_layer{lvl}")
        bottom = conv_train(up, x, rb[f"spconv{lvl}_inv"] if lvl != 1 else rb["subm1"])
    return bottom


def sigmoid_focal_loss(logits: torch.Tensor, labels: torch.Tensor, gamma: float = 2.0, alpha: float = 0.25) -> torch.Tensor:
    """mmdet FocalLoss(use_sigmoid=True, gamma 2, alpha 0.25): per-class binary focal terms, mean over points."""
    c = logits.size(1)   # the head's logits include the background column (num_classes + 1, segmentation_head.py:58-60)
    target = F.one_hot(labels.clamp(max=c - 1), c).to(logits.dtype)
    p = torch.sigmoid(logits)
    pt = (1 - p) * target + p * (1 - target)
    w = (alpha * target + (1 - alpha) * (1 - target)) * pt.pow(gamma)
    return (F.binary_cross_entropy_with_logits(logits, target, reduction="none") * w).sum() / max(logits.size(0), 1)


class SegmentorTrainer:
    """One training step of the segmentation stage on one frame per rank."""

    def __init__(self, model, lr: float = 1e-3, weight_decay: float = 0.01, bucket_mb: float = 25.0):
        self.model = model
        mods = [model.voxel_encoder, model.backbone_unet, model.segmentation_head]
        self.params: List[nn.Parameter] = [p for m in mods for p in m.parameters()]
        for p in model.parameters():
            p.requires_grad_(False)
        for p in self.params:
            p.requires_grad_(True)
        for m in mods:
            m.train()
        self.opt = torch.optim.AdamW(self.params, lr=lr, weight_decay=weight_decay)   # FSF_nuScenes_config.py: AdamW
        self.reducer = BucketedReducer(self.params, bucket_mb) if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1 else None

    def forward_loss(self, points: torch.Tensor, labels: torch.Tensor, vote_targets: torch.Tensor) -> Tuple[torch.Tensor, Dict[str, torch.Tensor]]:
        model, cfg = self.model, self.model.cfg
        P = model.point_dim
        pts = points[:, :P].contiguous()
        with torch.no_grad():
            coors3 = ops.voxelize(points, cfg["seg_voxel_size"], cfg["point_cloud_range"], floor_mode=0)
            coors4 = F.pad(coors3, (1, 0), value=0)
            plan = M.ScatterPlan(coors4, lo=[0, 0, 0, 0], ext=[1] + list(cfg["sparse_shape"]), want_index=True)
            rb, _ = model.backbone_unet.build_rulebooks(plan.new_coors, plan.index, 1)
        voxel = vfe_train(model.voxel_encoder, pts, coors4, plan)
        x = unet_train(model.backbone_unet, voxel, rb)
        # Voxel2PointScatterNeck (necks/voxel2point_neck.py:42-67): voxel feature of the point's voxel | xyz - voxel centre
        with torch.no_grad():
            vs = torch.tensor(cfg["seg_voxel_size"], device=pts.device)
            lo = torch.tensor(cfg["point_cloud_range"][:3], device=pts.device)
            centre = (coors3[:, [2, 1, 0]].float() + 0.5) * vs + lo
            local = pts[:, :3] - centre
        feats = torch.cat([x[plan.unq_inv], local], dim=1)
        head = model.segmentation_head
        h = mlp_train(head.pre_seg_conv, feats) if head.pre_seg_conv is not None else feats
        logits, votes = linear_train(head.conv_seg, h), linear_train(head.voting, h)
        loss_sem = sigmoid_focal_loss(logits, labels)
        # votes: v * |v| decoding (segmentation_head.py:262-266) of the labelled class' three channels, L1 on foreground points
        fg = labels < (head.num_classes - 1)
        if bool(fg.any()):
            idx = labels[fg].clamp(max=head.num_classes - 1)
            v = votes[fg].view(-1, head.num_classes, 3)[torch.arange(int(fg.sum()), device=pts.device), idx]
            loss_vote = F.l1_loss(v * v.abs(), vote_targets[fg])
        else:
            loss_vote = votes.sum() * 0
        return loss_sem + loss_vote, dict(loss_sem=loss_sem.detach(), loss_vote=loss_vote.detach())

    def step(self, points: torch.Tensor, labels: torch.Tensor, vote_targets: torch.Tensor) -> Dict[str, torch.Tensor]:
        self.opt.zero_grad(set_to_none=True)
        if self.reducer is not None:
            self.reducer.begin()
        loss, parts = self.forward_loss(points, labels, vote_targets)
        loss.backward()
        if self.reducer is not None:
            self.reducer.finish()
        self.opt.step()
        for m in (self.model.voxel_encoder, self.model.backbone_unet, self.model.segmentation_head):   # packed inference weights are stale
            for sub in m.modules():
                if hasattr(sub, "refresh"):
                    sub.refresh()
        parts["loss"] = loss.detach()
        return parts


# ------------------------------------------------------------------------------------------------
# bucketed, overlapped gradient all-reduce
# ------------------------------------------------------------------------------------------------
class BucketedReducer:
    """Flat fp32 buckets filled in reverse parameter order (the order backward produces gradients); a full bucket is all-reduced
    asynchronously on a side stream while backward goes on.  `begin()` before the forward pass, `finish()` after backward."""

    def __init__(self, params: List[nn.Parameter], bucket_mb: float = 25.0, group=None):
        self.params = list(params)
        self.group = group
        self.world = dist.get_world_size(group)
        cap = max(int(bucket_mb * (1 << 20) / 4), 1)
        self.buckets: List[dict] = []
        cur, size = [], 0
        for p in reversed(self.params):
            if cur and size + p.numel() > cap:
                self.buckets.append(dict(params=cur, numel=size))
                cur, size = [], 0
            cur.append(p)
            size += p.numel()
        if cur:
            self.buckets.append(dict(params=cur, numel=size))
        dev = self.params[0].device
        for b in self.buckets:
            b["flat"] = torch.zeros(b["numel"], dtype=torch.float32, device=dev)
            off = 0
            b["views"] = []
            for p in b["params"]:
                b["views"].append(b["flat"][off:off + p.numel()].view_as(p))
                off += p.numel()
        self.where = {id(p): (bi, pi) for bi, b in enumerate(self.buckets) for pi, p in enumerate(b["params"])}
        self.stream = torch.cuda.Stream(device=dev) if dev.type == "cuda" else None
        self.exposed_ms = 0.0
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in self.params]
        self.begin()

    def begin(self):
        for b in self.buckets:
            b["pending"] = len(b["params"])
            b["work"] = None
            b["launched"] = False

    def _on_grad(self, p: nn.Parameter):
        bi, pi = self.where[id(p)]
        b = self.buckets[bi]
        b["views"][pi].copy_(p.grad)
        b["pending"] -= 1
        if b["pending"] == 0:
            self._launch(b)

    def _launch(self, b: dict):
        b["launched"] = True
        if self.stream is not None:
            self.stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.stream):
                b["work"] = dist.all_reduce(b["flat"], op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        else:
            b["work"] = dist.all_reduce(b["flat"], op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def finish(self):
        """Parameters that received no gradient this step count as zeros (every rank must reduce every bucket)."""
        for b in self.buckets:
            if not b["launched"]:
                for pi, p in enumerate(b["params"]):
                    if p.grad is None:
                        b["views"][pi].zero_()
                self._launch(b)
        ev0 = ev1 = None
        if self.stream is not None:
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
        for b in self.buckets:
            b["work"].wait()
        if self.stream is not None:
            torch.cuda.current_stream().wait_stream(self.stream)
            ev1.record()
        for b in self.buckets:
            b["flat"].div_(self.world)
            for pi, p in enumerate(b["params"]):
                if p.grad is None:
                    p.grad = b["views"][pi].clone()
                else:
                    p.grad.copy_(b["views"][pi])
        if ev0 is not None:
            self._events = (ev0, ev1)

    def exposed_time_ms(self) -> float:
        """Device time between the end of backward and the last bucket's arrival (call after a synchronize)."""
        ev = getattr(self, "_events", None)
        return float(ev[0].elapsed_time(ev[1])) if ev else 0.0
