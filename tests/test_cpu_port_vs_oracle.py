"""Cross-validation of the two CPU restatements: the torch-CPU port used as the timed baseline
(oracle/fsf_torch_cpu.py) against the numpy oracle blocks, on a small frame."""
import numpy as np
import torch

from fullysparsefusion_b200 import fsf as FSFM
from fullysparsefusion_b200 import synth
from oracle import fsf_oracle as O
from oracle import fsf_oracle_frame as OF
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
