"""The reference plugin imported UNCHANGED with the B200 shims in place of torch_scatter / ingroup_indices / torchex /
dynamic_point_pool_ext / mmdet3d.ops.  Needs /root/reference (build container only; the GPU box has no reference
tree, which is why nothing here runs a kernel): mmcv / mmdet / the rest of mmdet3d are absent from this image and are served by
the inert stub finder of tools/make_golden.py, everything the hot path binds at import time is the shim."""
import os
import subprocess
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"

CHILD = r"""
import importlib, sys, types
sys.path.insert(0, %(repo)r)
sys.path.insert(0, %(repo)r + "/tools")
import fullysparsefusion_b200.shims as shims
from fullysparsefusion_b200.shims import mmdet3d_ops
import make_golden as G
G.STUB_ROOTS = tuple(r for r in G.STUB_ROOTS if r not in ("torch_scatter", "ingroup_indices", "dynamic_point_pool_ext", "torchex"))
sys.meta_path.insert(0, G._Finder())      # mmcv / mmdet / mmdet3d / mmseg: inert stand-ins (absent from this image)
shims.install(force=True)                 # the four extension modules + the three mmdet3d.ops names
assert sys.modules["mmdet3d.ops"].Voxelization is mmdet3d_ops.Voxelization
reg = {}
def build_voxel_encoder(cfg):
    cfg = dict(cfg); t = cfg.pop("type")
    cls = {t_: c for r, t_, c in shims.registry_table() if r == "VOXEL_ENCODERS"}[t]
    reg.setdefault(t, 0); reg[t] += 1
    return cls(**cfg)
G._SPECIAL[("mmdet3d.models", "builder")] = types.SimpleNamespace(build_voxel_encoder=build_voxel_encoder)
G._SPECIAL[("mmdet3d.models.builder", "build_voxel_encoder")] = build_voxel_encoder
del G._SPECIAL[("torch_scatter", "scatter_max")], G._SPECIAL[("torch_scatter", "scatter")]
mods = G.import_reference()
sst, fsd = mods["sst_ops"], mods["fsd"]
pool = importlib.import_module("projects.mmdet3d_plugin.ops.dynamic_point_pool_op")
from fullysparsefusion_b200.shims import torch_scatter as ts, ingroup_indices as ii, torchex as tx, dynamic_point_pool_ext as dp
assert sst.torch_scatter is ts and sst.ingroup_indices is ii                       # ops/sst_ops.py:6,239
assert sst.spconv is mmdet3d_ops.spconv                                             # ops/sst_ops.py:5
assert fsd.Voxelization is mmdet3d_ops.Voxelization and fsd.furthest_point_sample is mmdet3d_ops.furthest_point_sample   # :13
assert fsd.cc_gpu is tx.connected_components                                        # single_stage_fsd.py:20-23
assert pool.dynamic_point_pool_ext is dp                                            # ops/dynamic_point_pool_op.py:5
shims.patch_ccl(fsd)
assert fsd.find_connected_componets is tx.find_connected_componets
# the reference's own module classes build the registry types from the repo's classes with the kwargs they pass themselves
sir_mod = importlib.import_module("projects.mmdet3d_plugin.models.backbones.sir")
head_mod = importlib.import_module("projects.mmdet3d_plugin.models.roi_heads.bbox_heads.fsd_bbox_head")
sir_mod.SIR(num_blocks=3, in_channels=[16, 35, 35], feat_channels=[[32, 32]] * 3, rel_mlp_hidden_dims=[[16, 32]] * 3, norm_cfg=dict(type="LN", eps=1e-3),
            xyz_normalizer=[20, 20, 4], act="gelu")
head_mod.FullySparseBboxHead(num_classes=10, num_blocks=3, in_channels=[32, 48, 48], feat_channels=[[32, 32]] * 3, with_distance=False,
                             with_cluster_center=False, with_rel_mlp=True, rel_mlp_hidden_dims=[[16, 32]] * 3, rel_mlp_in_channels=[13] * 3,
                             reg_mlp=None, cls_mlp=None)
assert reg == {"SIRLayer": 3, "DynamicClusterVFE": 3}, reg
# shims refuse CPU tensors (no CPU fallback behind the reference's call sites)
import torch
try:
    sst.scatter_v2(torch.randn(8, 4), torch.zeros(8, 3, dtype=torch.long), "max")
except Exception as e:
    assert "CUDA" in str(e) or "cuda" in str(e), e
else:
    raise AssertionError("scatter_v2 on CPU tensors must raise")
print("PLUGIN OK")
"""


@pytest.mark.skipif(not os.path.isdir(REF), reason="needs /root/reference (build container only)")
def test_plugin_imports_bind_to_the_shims():
    r = subprocess.run([sys.executable, "-c", CHILD % dict(repo=REPO)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "PLUGIN OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
