"""Training step of the segmentation stage (fullysparsefusion_b200/train.py) on device.

* the unfused training forward (autograd.sparse_conv + BatchNorm modules + torch.cat) equals the fused inference forward of the
  same parameters when the norms are in eval mode — pins the U-Net topology of the trainable path (concatenation order,
  reduce_channel, residuals, inverse-convolution rulebooks);
* gradients reach every parameter, are finite, and a few AdamW steps on one frame lower the loss;
* the packed inference weights are refreshed after a step (the detector sees the trained parameters)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import bench
from fullysparsefusion_b200 import modules as M
from fullysparsefusion_b200 import ops, synth, train

pytestmark = pytest.mark.gpu


def _frame(cuda, n=20000, seed=3):
    pts = torch.from_numpy(synth.ring_points(n, sweeps=1, seed=seed)).to(cuda)
    labels, votes = bench.synth_labels(pts)
    return pts, labels, votes


def test_training_forward_equals_inference_forward_in_eval_mode(cuda):
    torch.manual_seed(0)
    model = bench.make_model().to(cuda).eval()
    for m in model.modules():                       # non-trivial running statistics
        if isinstance(m, torch.nn.BatchNorm1d):
            m.running_mean.normal_(0, 0.1)
            m.running_var.uniform_(0.5, 1.5)
    for m in model.modules():
        if hasattr(m, "refresh"):
            m.refresh()
    pts, _, _ = _frame(cuda)
    cfg, P = model.cfg, model.point_dim
    with torch.no_grad():
        coors3 = ops.voxelize(pts, cfg["seg_voxel_size"], cfg["point_cloud_range"], floor_mode=0)
        coors4 = F.pad(coors3, (1, 0), value=0)
        plan = M.ScatterPlan(coors4, lo=[0, 0, 0, 0], ext=[1] + list(cfg["sparse_shape"]), want_index=True)
        pts5 = pts[:, :P].contiguous()
        want_vox, voxel_coors, _ = model.voxel_encoder(pts5, coors4, return_inv=True, plan=plan)
        rb, _ = model.backbone_unet.build_rulebooks(voxel_coors, plan.index, 1)
        want = model.backbone_unet(dict(voxel_feats=want_vox, voxel_coors=voxel_coors), rulebooks=rb)[0]["voxel_feats"]
        got_vox = train.vfe_train(model.voxel_encoder, pts5, coors4, plan)
        got = train.unet_train(model.backbone_unet, got_vox, rb)
    scale = float(want.abs().max())
    assert float((got_vox - want_vox).abs().max()) <= 1e-4 * float(want_vox.abs().max())
    assert float((got - want).abs().max()) <= 2e-4 * scale, float((got - want).abs().max()) / scale


def test_training_step_updates_every_parameter_and_lowers_the_loss(cuda):
    torch.manual_seed(0)
    model = bench.make_model().to(cuda)
    trainer = train.SegmentorTrainer(model, lr=2e-3)
    pts, labels, votes = _frame(cuda)
    before = [p.detach().clone() for p in trainer.params]
    losses = []
    for _ in range(6):
        out = trainer.step(pts, labels, votes)
        losses.append(float(out["loss"]))
        assert np.isfinite(losses[-1])
    for p in trainer.params:
        assert p.grad is not None and bool(torch.isfinite(p.grad).all())
    changed = sum(int(not torch.equal(a, b.detach())) for a, b in zip(before, trainer.params))
    assert changed == len(trainer.params), (changed, len(trainer.params))
    assert losses[-1] < losses[0], losses
    # the inference path sees the trained weights (packed copies were dropped by the step)
    model.eval()
    with torch.no_grad():
        x = model.backbone_unet.conv_input.weight
        pack = ops.gemm_prepack(x.detach().float())
        assert model.backbone_unet.conv_input._pack is None or model.backbone_unet.conv_input._pack[0].data.shape == pack.data.shape


def test_submanifold_input_gradient_reuses_the_rulebook(cuda):
    """A submanifold rulebook transposed is itself with the offsets reversed: the input gradient through that shortcut (row order
    and row-ordered table reused) equals the one through the explicit transposition."""
    from fullysparsefusion_b200 import autograd as AG

    pts = torch.from_numpy(synth.ring_points(120000, sweeps=4, seed=5)).to(cuda)
    coors4 = F.pad(ops.voxelize(pts, synth.NUSC_VOXEL, synth.NUSC_RANGE, floor_mode=0), (1, 0), value=0)
    plan = M.ScatterPlan(coors4, lo=[0, 0, 0, 0], ext=[1, 40, 512, 512], want_index=True)
    rb = M.Rulebook(ops.conv_rulebook(plan.new_coors, plan.index, 3, 1, 1))
    assert rb.order is not None and rb.nbr_ro is not None      # big enough for the row order
    g = torch.Generator(device=cuda).manual_seed(1)
    a = torch.randn(plan.m, 64, device=cuda, generator=g)
    w = torch.randn(27, 32, 64, device=cuda, generator=g) / 40
    coef = torch.randn(plan.m, 32, device=cuda, generator=g)
    grads = []
    for sym in (True, False):
        a1, w1 = a.clone().requires_grad_(True), w.clone().requires_grad_(True)
        y = AG.sparse_conv(a1, w1, rb.nbr, rb.order, rb.nbr_ro, symmetric=sym)
        (y * coef).sum().backward()
        grads.append((y.detach(), a1.grad, w1.grad))
    assert torch.equal(grads[0][0], grads[1][0])
    scale = float(grads[1][1].abs().max())      # two summation orders of the same fp32 sums
    assert float((grads[0][1] - grads[1][1]).abs().max()) <= 2e-5 * scale, float((grads[0][1] - grads[1][1]).abs().max()) / scale
    assert torch.equal(grads[0][2], grads[1][2])
