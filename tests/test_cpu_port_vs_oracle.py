"""Cross-validation of the two CPU restatements: the torch-CPU port used as the timed baseline
(oracle/fsf_torch_cpu.py) against the numpy oracle blocks, on a small frame."""
import numpy as np
import torch

from fullysparsefusion_b200 import fsf as FSFM
from fullysparsefusion_b200 import synth
from oracle import fsf_oracle as O
from oracle import fsf_oracle_models as OM
from oracle import fsf_torch_cpu as P


def test_torch_cpu_port_matches_numpy_oracle():
    n, H, W = 900, 90, 160
    pts = synth.ring_points(n, sweeps=1, seed=31)
    mask = synth.mask_planes(6, 10, H, W, seed=31, overlap=True)
    anno = synth.mask_anno(mask, seed=31)
    l2i = synth.lidar2img(6, H, W)
    torch.manual_seed(0)
    model = FSFM.FSF()
    with torch.no_grad():
        model.segmentor_updated_mlp[-1].weight.normal_(0, 0.05)
        model.segmentation_head.conv_seg.bias.copy_(torch.linspace(-1.0, 1.0, 11))
    cpu = P.CpuFSF(model)
    with torch.no_grad():
        stages, st = cpu.stages(torch.from_numpy(pts), torch.from_numpy(mask), torch.from_numpy(anno), torch.from_numpy(l2i))
        for _, fn in stages:
            fn()
    sd = {k: v.detach().numpy() for k, v in model.state_dict().items()}
    sub = lambda p: {k[len(p):]: v for k, v in sd.items() if k.startswith(p)}
    cfg = FSFM.NUSC
    coors = np.concatenate([np.zeros((n, 1), np.int64), O.voxelize(pts, cfg["seg_voxel_size"], cfg["point_cloud_range"], 1).astype(np.int64)], 1)
    vf, vc, inv = OM.dynamic_scatter_vfe(pts[:, :5], coors, sub("voxel_encoder."), cfg["seg_voxel_size"], cfg["point_cloud_range"])
    net = model.backbone_unet
    want, _, _ = OM.simple_sparse_unet(vf, vc, sub("backbone_unet."), cfg["sparse_shape"], net.encoder_channels, net.encoder_paddings,
                                       ((512, 512, 256), (256, 256, 128), (128, 128, 128), (128, 128, 128), (128, 128, 128)))
    np.testing.assert_allclose(st["voxel_feats"].numpy(), want, rtol=2e-3, atol=2e-3)       # fp32 CPU vs float64 oracle, 34 layers
    ids = O.points_in_mask(pts[:, 5:8], mask, l2i)
    assert np.mean(np.any(st["ids"].numpy() != ids, axis=(1, 2))) <= 5e-3                   # ATen grid_sample vs oracle: texel-boundary flips only
    logits, votes = OM.vote_seg_head(st["seg_feats"].numpy(), sub("segmentation_head."))
    np.testing.assert_allclose(st["seg_logits"].numpy(), logits, rtol=1e-3, atol=1e-4)
    assert st["obj_feats"].shape[1] == 1024 and st["frustum_obj_feats"].shape[1] == 896


def test_torch_cpu_port_refine_and_boxes_match_numpy_oracle():
    """The port's optional refine / boxes stages (not in the timed scope) against the numpy oracle blocks."""
    n, H, W = 900, 90, 160
    pts = synth.ring_points(n, sweeps=1, seed=32)
    mask = synth.mask_planes(6, 10, H, W, seed=32, overlap=True)
    anno = synth.mask_anno(mask, seed=32)
    l2i = synth.lidar2img(6, H, W)
    torch.manual_seed(1)
    model = FSFM.FSF()
    with torch.no_grad():
        model.segmentor_updated_mlp[-1].weight.normal_(0, 0.05)
        model.segmentation_head.conv_seg.bias.copy_(torch.linspace(-1.0, 1.0, 11))
        # small boxes decode to exp(~0) - 1e-6 ~ 1 m: make the heads emit a spread of sizes so that rois catch points and overlap
        for h in (model.frustum_obj_head, model.bbox_head):
            h.task_heads[0].dim[-1].bias.copy_(torch.tensor([0.6, 1.4, 0.5]))
    cpu = P.CpuFSF(model)
    with torch.no_grad():
        stages, st = cpu.stages(torch.from_numpy(pts), torch.from_numpy(mask), torch.from_numpy(anno), torch.from_numpy(l2i))
        for _, fn in stages + cpu.extra_stages:
            fn()
    sd = {k: v.detach().numpy() for k, v in model.state_dict().items()}
    sub = lambda p: {k[len(p):]: v for k, v in sd.items() if k.startswith(p)}
    rois = st["refine0_rois"].numpy()
    K = rois.shape[0]
    assert K > 0
    reg = np.concatenate([st["frustum_out"][1].numpy(), st["fsd_out"][1].numpy()])
    ctr = np.concatenate([st["frustum_centers"].numpy(), st["fsd_centers"].numpy()])
    np.testing.assert_allclose(rois, O.decode_boxes(reg, ctr), rtol=1e-5, atol=1e-5)
    wp, wr, wf, amb = O.dynamic_point_pool(rois[:, 1:8], pts[:, :3], [1.0, 1.0, 1.0], 512, 10 ** 7, margin=1e-4)
    gp, gr = st["refine0_pts_inds"].numpy(), st["refine0_roi_inds"].numpy()
    assert len(wp) > 50
    assert (set(zip(gr.tolist(), gp.tolist())) ^ set(zip(wr.tolist(), wp.tolist()))) <= amb
    if np.array_equal(gp, wp) and np.array_equal(gr, wr):       # no face-ambiguous pair in this frame: the head sees the same rows
        with torch.no_grad():
            img = P.seq(model.refine_img_mlp[0], st["img_scores"][torch.from_numpy(wp)]).numpy()
        feats = np.concatenate([st["seg_feats"].numpy()[wp], img], 1)
        want, wmask = OM.fully_sparse_bbox_head(pts[wp, :5], feats, wf[:, 3:6], wf[:, 6:12], wf[:, 12], wr, rois, sub("refine_sir_layers.0."), 3)
        np.testing.assert_allclose(st["refine0_lidar_feat"].numpy(), want, rtol=2e-3, atol=2e-3)
        assert np.array_equal(np.abs(st["refine0_lidar_feat"].numpy()).sum(1) > 0, wmask)
    boxes = O.decode_boxes(st["refine0_reg"].numpy(), st["refine0_centers"].numpy())[:, 1:]
    wb, ws, wl, wrow = O.multiclass_nms(boxes, st["refine0_cls"].numpy(), 0.01, 0.35, 500)
    assert np.array_equal(st["det_rows"].numpy(), wrow) and np.array_equal(st["det_labels"].numpy(), wl)
    np.testing.assert_allclose(st["det_scores"].numpy(), ws, rtol=1e-6)


def test_reference_checkpoint_key_mapping():
    """FSF.load_reference_state_dict: the reference nests the segmentor's modules (`segmentor.*`) and spconv stores kernels 5-d."""
    torch.manual_seed(3)
    src = FSFM.FSF()
    ref_sd = {}
    for k, v in src.state_dict().items():
        for a, b in FSFM.FSF.REFERENCE_PREFIXES:
            if k.startswith(b):
                k = a + k[len(b):]
                break
        ref_sd[k] = v.clone()
    n5 = 0
    for k in list(ref_sd):
        v = ref_sd[k]
        if v.dim() == 3 and v.size(0) == 27:                        # conv kernels, alternately in the two spconv layouts
            koff, cout, cin = v.shape
            if n5 % 2 == 0:
                ref_sd[k] = v.reshape(3, 3, 3, cout, cin).permute(0, 1, 2, 4, 3).contiguous()       # spconv 1.x [kz,ky,kx,Cin,Cout]
            else:
                ref_sd[k] = v.reshape(3, 3, 3, cout, cin).permute(3, 0, 1, 2, 4).contiguous()       # spconv 2.x [Cout,kz,ky,kx,Cin]
            n5 += 1
    assert n5 > 20 and any(k.startswith("segmentor.backbone.") for k in ref_sd)
    torch.manual_seed(4)
    dst = FSFM.FSF()
    missing, unexpected = dst.load_reference_state_dict(ref_sd, strict=True)
    assert not missing and not unexpected
    for (k, a), (_, b) in zip(src.state_dict().items(), dst.state_dict().items()):
        assert torch.equal(a, b), k
    ref_sd["segmentor.backbone.not_a_key"] = torch.zeros(1)
    missing, unexpected = dst.load_reference_state_dict(ref_sd, strict=False)
    assert unexpected == ["backbone_unet.not_a_key"] and not missing
