"""Drop-in modules under the names the reference plugin imports.

    import fullysparsefusion_b200.shims as shims
    shims.install()          # registers torch_scatter, ingroup_indices, torchex, dynamic_point_pool_ext in sys.modules
    import projects.mmdet3d_plugin   # the plugin's `import torch_scatter` etc. now bind to the B200 path

Every function is a thin wrapper over the C-ABI ops (no torch arithmetic on the hot path, no CPU
fallback: CPU tensors raise).  torch_scatter carries autograd (torch_scatter's own gradient rules); the other shims return
index tensors, which are non-differentiable upstream too.  The rest of training backward (sparse-conv dgrad / wgrad, SyncBN
statistics) is SURVEY.md §8f row 4 and not built.
"""
from __future__ import annotations

import sys

from . import dynamic_point_pool_ext, ingroup_indices, mmdet3d_ops, torch_scatter, torchex  # noqa: F401


def install(force: bool = False) -> None:
    """Register the shims under the import names used by projects/mmdet3d_plugin
    (ops/sst_ops.py:5-6,239; ops/dynamic_point_pool_op.py:5; models/detectors/single_stage_fsd.py:13,20-23)."""
    for name, mod in (("torch_scatter", torch_scatter), ("ingroup_indices", ingroup_indices), ("torchex", torchex),
                      ("dynamic_point_pool_ext", dynamic_point_pool_ext)):
        if force or name not in sys.modules:
            sys.modules[name] = mod
    # mmdet3d.ops (Voxelization, furthest_point_sample, spconv: single_stage_fsd.py:13, sst_ops.py:5): with an mmdet3d package
    # importable the three names are patched INTO its `ops` package (the plugin also imports mmdet3d.ops.roiaware_pool3d /
    # iou3d submodules at load time, which stay the fork's); without mmdet3d the shim module stands in under that name
    try:
        import importlib

        real = importlib.import_module("mmdet3d.ops")
    except Exception:
        real = None
    if real is not None and real is not mmdet3d_ops:
        for attr in ("Voxelization", "furthest_point_sample", "spconv"):
            setattr(real, attr, getattr(mmdet3d_ops, attr))
    else:
        sys.modules["mmdet3d.ops"] = mmdet3d_ops


def patch_ccl(single_stage_fsd_module) -> None:
    """Route the stock config's live CCL path (scipy on the host, single_stage_fsd.py:45-82) to the
    CUDA kernel by replacing the module-level functions ClusterAssigner looks up at call time
    (dispatch at single_stage_fsd.py:971-977)."""
    single_stage_fsd_module.find_connected_componets = torchex.find_connected_componets
    single_stage_fsd_module.find_connected_componets_single_batch = torchex.find_connected_componets_single_batch
    single_stage_fsd_module.cc_gpu = torchex.connected_components


def registry_table():
    """(registry attribute in mmdet3d / mmdet, type string the configs use, class here) — the rows INTEGRATION.md section 3 lists."""
    from .. import fsf, loading, modules

    return [
        ("VOXEL_ENCODERS", "DynamicScatterVFE", modules.DynamicScatterVFE),     # FSF_nuScenes_config.py:42-52
        ("VOXEL_ENCODERS", "SIRLayer", modules.SIRLayer),                       # sir.py:61
        ("VOXEL_ENCODERS", "DynamicClusterVFE", modules.DynamicClusterVFE),     # fsd_bbox_head.py:62-87
        ("NORM_LAYERS", "naiveSyncBN1d", modules.naiveSyncBN1d),                # mmcv.cnn NORM_LAYERS; sst_ops.py:814
        ("BACKBONES", "SimpleSparseUNet", modules.SimpleSparseUNet),            # config :58-70
        ("BACKBONES", "SIR", modules.SIR),                                      # config :113-124, :201-212
        ("NECKS", "Voxel2PointScatterNeck", modules.Voxel2PointScatterNeck),    # config :72-76
        ("HEADS", "VoteSegHead", modules.VoteSegHead),                          # config :78-95
        ("HEADS", "SparseClusterHeadV2", fsf.SparseClusterHeadV2),              # config :126-160
        ("HEADS", "FrustumClusterHead", fsf.SparseClusterHeadV2),               # same forward (inherits), config :214-273
        ("HEADS", "FSDSeparateHead", fsf.FSDSeparateHead),
        ("HEADS", "FullySparseBboxHead", modules.FullySparseBboxHead),          # config :296-320
        ("ROI_EXTRACTORS", "DynamicPointROIExtractor", modules.DynamicPointROIExtractor),   # config :290-294
        ("PIPELINES", "LoadMaskFromFiles", loading.LoadMaskFromFiles),
        ("PIPELINES", "SaveNoAugPoints", loading.SaveNoAugPoints),
    ]


def register(registries, force: bool = True):
    """Register the classes above in mmcv registries.  `registries`: a module or mapping that exposes the registry objects by the
    attribute names of `registry_table()` (e.g. `mmdet3d.models.builder` for the model registries, `mmdet.datasets.builder` for
    PIPELINES); registries it does not have are skipped.  Uses mmcv's `Registry.register_module(name=, force=, module=)`.
    Returns the (registry, type) pairs that were registered."""
    done = []
    for reg_name, type_name, cls in registry_table():
        reg = registries.get(reg_name) if isinstance(registries, dict) else getattr(registries, reg_name, None)
        if reg is None:
            continue
        reg.register_module(name=type_name, force=force, module=cls)
        done.append((reg_name, type_name))
    return done
