"""The C-ABI library builds, loads and exports every symbol include/fsf_b200.h declares (no GPU)."""
import ctypes
import re
from pathlib import Path

import pytest

REPO = Path(__file__).resolve().parents[1]


def _declared():
    text = (REPO / "include" / "fsf_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fsfb_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_symbols():
    names = _declared()
    assert "fsfb_segment_reduce" in names and "fsfb_project_sample" in names and len(names) >= 12


def test_library_exports_every_declared_symbol():
    from fullysparsefusion_b200 import _capi

    lib = _capi.load()
    raw = ctypes.CDLL(str(_capi.LIB_PATH))
    for name in _declared():
        assert hasattr(raw, name), f"{name} declared in fsf_b200.h but not exported"
        assert name in _capi.SIGNATURES, f"{name} has no ctypes signature in _capi.py"
    for name in _capi.SIGNATURES:
        assert name in _declared(), f"{name} bound in _capi.py but not declared in the header"
    assert lib.fsfb_version() >= 1
    assert lib.fsfb_launch_count() >= 0


def test_bad_arguments_fail_without_gpu():
    from fullysparsefusion_b200 import _capi

    lib = _capi.load()
    need = ctypes.c_size_t(0)
    assert lib.fsfb_rank_workspace_bytes(10, 1 << 40, ctypes.byref(need)) == _capi.ERR_BADARG
    assert b"2^32" in lib.fsfb_last_error()
    assert lib.fsfb_voxelize(None, -1, 3, None, None, None, 0, 0, 1, None, None) == _capi.ERR_BADARG
    assert lib.fsfb_csr_workspace_bytes(100, 10, ctypes.byref(need)) == 0 and need.value > 0


def test_ops_refuse_cpu_tensors():
    import torch

    from fullysparsefusion_b200 import _capi, ops

    with pytest.raises(_capi.FsfbError):
        ops.voxelize(torch.zeros(4, 5), [0.2] * 3, [-1, -1, -1, 1, 1, 1])
    with pytest.raises(_capi.FsfbError):
        ops.unique_rows(torch.zeros(4, 3, dtype=torch.int64))


def test_prepack_size_with_fp16_blocks():
    """FSFB_GEMM_F16=1 (experimental, read once per process) appends the fp16-split blocks: half the bytes of the tf32 blocks."""
    import os
    import subprocess
    import sys

    code = ("import ctypes; from fullysparsefusion_b200 import _capi; lib = _capi.load(); n = ctypes.c_size_t(0); out = []\n"
            "for k, ci, co in [(27, 128, 128), (1, 131, 11), (27, 64, 512), (1, 768, 1024)]:\n"
            "    assert lib.fsfb_gemm_prepack_bytes(k, ci, co, ctypes.byref(n)) == 0; out.append(n.value)\n"
            "print(*out)")
    sizes = {}
    for flag in ("0", "1"):
        r = subprocess.run([sys.executable, "-c", code], cwd=str(REPO), env=dict(os.environ, FSFB_GEMM_F16=flag), capture_output=True,
                           text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        sizes[flag] = [int(v) for v in r.stdout.split()]
    assert all(b * 2 == a * 3 for a, b in zip(sizes["0"], sizes["1"])), sizes


def test_registry_registration_with_a_stand_in_registry():
    """shims.register drives mmcv's Registry API; mmcv is not installed here, so a registry with the same method stands in."""
    from fullysparsefusion_b200 import shims

    class Registry:
        def __init__(self):
            self.module_dict = {}

        def register_module(self, name=None, force=False, module=None):
            assert module is not None and (force or name not in self.module_dict)
            self.module_dict[name] = module

    regs = {n: Registry() for n in ("VOXEL_ENCODERS", "BACKBONES", "HEADS", "PIPELINES")}       # NECKS / ROI_EXTRACTORS absent: skipped
    done = shims.register(regs)
    assert ("HEADS", "VoteSegHead") in done and ("NECKS", "Voxel2PointScatterNeck") not in done
    assert set(regs["VOXEL_ENCODERS"].module_dict) == {"DynamicScatterVFE", "SIRLayer", "DynamicClusterVFE"}
    assert regs["PIPELINES"].module_dict["LoadMaskFromFiles"].__name__ == "LoadMaskFromFiles"
    assert len(done) == len([r for r in shims.registry_table() if r[0] in regs])
    shims.register(regs)        # force=True: a second registration replaces, as mmcv allows


@pytest.mark.parametrize("cfg_file", ["nuScenes/FSF_nuScenes_config.py", "Argoverse2/FSF_AV2_config.py"])
def test_stock_config_kwargs_construct_the_drop_in_classes(cfg_file):
    """Every model sub-config of the reference's stock configs whose `type` is in shims.registry_table() constructs the class here
    with its kwargs unchanged (build container only: reads the config from /root/reference)."""
    import copy
    import os

    path = os.path.join("/root/reference/projects/configs", cfg_file)
    if not os.path.exists(path):
        pytest.skip("reference configs are not on this box")
    from fullysparsefusion_b200 import shims

    ns = {}
    exec(compile(open(path).read(), path, "exec"), ns)
    table = {t: c for reg, t, c in shims.registry_table() if reg != "NORM_LAYERS"}   # a norm cfg is completed with num_features
    found = []

    def walk(d):
        if isinstance(d, dict):
            if d.get("type") in table and d.get("type") != "FSDSeparateHead":     # a partial spec its parent head completes
                found.append(d)
            for v in d.values():
                walk(v)
        elif isinstance(d, (list, tuple)):
            for v in d:
                walk(v)

    walk(ns["model"])
    assert len(found) >= 9
    for d in found:
        table[d["type"]](**{k: copy.deepcopy(v) for k, v in d.items() if k != "type"})
