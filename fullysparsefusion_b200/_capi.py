"""ctypes binding of libfsf_b200.so (the C-ABI declared in include/fsf_b200.h).

The product path has no CPU fallback: if the library is missing this module raises at import
of any op, and every op refuses non-CUDA tensors.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "_lib" / "libfsf_b200.so"

OK, ERR_BADARG, ERR_CAPACITY, ERR_CUDA = 0, -1, -2, -3
REDUCE_SUM, REDUCE_MEAN, REDUCE_MAX = 0, 1, 2
ACT_NONE, ACT_RELU, ACT_GELU = 0, 1, 2
NORM_NONE, NORM_LAYERNORM, NORM_AFFINE = 0, 1, 2

_p, _i, _i64, _sz, _f = C.c_void_p, C.c_int, C.c_int64, C.c_size_t, C.c_float
_psz = C.POINTER(C.c_size_t)

# name -> (restype, argtypes).  Keep in the same order as include/fsf_b200.h.
SIGNATURES = {
    "fsfb_version": (_i, []),
    "fsfb_last_error": (C.c_char_p, []),
    "fsfb_launch_count": (_i64, []),
    "fsfb_voxelize": (_i, [_p, _i64, _i64, _p, _p, _p, _i, _i, _i, _p, _p]),
    "fsfb_rows_minmax": (_i, [_p, _i, _i64, _i, _p, _p]),
    "fsfb_rank_workspace_bytes": (_i, [_i64, _i64, _psz]),
    "fsfb_rank_rows": (_i, [_p, _i, _i64, _i, _p, _p, _p, _sz, _p, _p, _p, _i64, _p, _p, _p, _p]),
    "fsfb_csr_workspace_bytes": (_i, [_i64, _i64, _psz]),
    "fsfb_csr_build": (_i, [_p, _i, _i64, _i64, _p, _p, _p, _p, _sz, _p]),
    "fsfb_segment_reduce_workspace_bytes": (_i, [_i64, _i, _i, _psz]),
    "fsfb_segment_reduce": (_i, [_p, _i64, _i, _i64, _p, _p, _p, _i64, _i, _p, _p, _p, _sz, _p]),
    "fsfb_gather_rows": (_i, [_p, _i64, _i, _i64, _p, _i, _i64, _f, _p, _i64, _p]),
    "fsfb_ingroup_workspace_bytes": (_i, [_i64, _i64, _psz]),
    "fsfb_ingroup_indices": (_i, [_p, _i64, _i64, _p, _p, _sz, _p]),
    "fsfb_project_sample": (_i, [_p, _i64, _i64, _p, _i, _p, _i, _i, _i, _i, _p, _p]),
    "fsfb_project_sample_select": (_i, [_p, _i64, _i64, _p, _i, _p, _i, _i, _i, _i, _p, _p, _p, _p, _p, _i, _i, _i, _p, _p]),
    "fsfb_project_sample_select_hwc": (_i, [_p, _i64, _i64, _p, _i, _p, _i, _i, _i, _p, _p, _p, _p, _p, _i, _i, _i, _p, _p]),
    "fsfb_gemm_prepack_bytes": (_i, [_i, _i, _i, _psz]),
    "fsfb_gemm_prepack": (_i, [_p, _i, _i, _i, _p, _p]),
    "fsfb_gather_gemm": (_i, [_p, _i64, _i, _i64, _p, _p, _i, _i64, _p, _i, _p, _i, _p, _p, _f, _p, _i64, _i, _p, _i64, _p]),
    "fsfb_gather_gemm_splitk_bytes": (_i, [_i64, _i, _i, _p]),
    "fsfb_gather_gemm_splitk": (_i, [_p, _i64, _i, _i64, _p, _p, _i, _i64, _p, _i, _p, _i, _p, _p, _f, _p, _i64, _i, _p, _i64, _i, _p, _sz, _p]),
    "fsfb_debug_gemm_timers": (_i, [_p]),
    "fsfb_debug_gemm_ss_timers": (_i, [_p]),
    "fsfb_gemm_f16_overflows": (_i, [_p]),
    "fsfb_split_rows": (_i, [_p, _i64, _i, _i64, _p, _p]),
    "fsfb_gather_gemm_split": (_i, [_p, _i64, _i, _p, _p, _i, _i64, _p, _i, _p, _i, _p, _p, _f, _p, _i64, _i, _p, _i64, _i, _p, _sz,
                                   _p, _p, _p, _p]),
    "fsfb_group_flags": (_i, [_p, _i64, _i64, _i, _p, _p, _p, _p]),
    "fsfb_group_split": (_i, [_p, _i64, _i64, _i, _p, _p, _p, _p]),
    "fsfb_group_voxelize": (_i, [_p, _i64, _p, _p, _p, _i, _p, _p]),
    "fsfb_group_keep": (_i, [_p, _p, _p, _i64, _i, _i, _p, _p, _p]),
    "fsfb_group_relabel": (_i, [_p, _p, _i64, _i64, _i, _p, _i64, _p, _p, _p]),
    "fsfb_connected_components_groups": (_i, [_p, _i64, _i64, _p, _p, _i, _p, _p, _p, _sz, _p]),
    "fsfb_nms_flags": (_i, [_p, _i64, _i, _i64, _i, _f, _p, _p, _p, _p]),
    "fsfb_nms_workspace_bytes": (_i, [_i64, _i, _p]),
    "fsfb_nms_suppress": (_i, [_p, _i64, _i64, _p, _i, _p, _i64, _p, _i, _f, _p, _p, _sz, _p]),
    "fsfb_nms_emit": (_i, [_p, _i64, _i, _p, _i64, _i64, _i, _i64, _p, _p, _p, _p, _p, _sz, _p]),
    "fsfb_decode_boxes": (_i, [_p, _i64, _i, _i64, _p, _i64, _p, _i64, _p, _p]),
    "fsfb_dynamic_point_pool_workspace_bytes": (_i, [_i64, _i, _p]),
    "fsfb_permute_rulebook": (_i, [_p, _i, _i64, _p, _p, _p]),
    "fsfb_conv_wgrad_workspace_bytes": (_i, [_i64, _i, _i, _i, _p]),
    "fsfb_conv_wgrad": (_i, [_p, _i64, _i, _i64, _p, _i64, _i, _i64, _p, _i, _p, _p, _sz, _p]),
    "fsfb_dynamic_point_pool": (_i, [_p, _i64, _p, _i64, _i64, _p, _i, _i64, _p, _p, _p, _p, _p, _sz, _p]),
    "fsfb_gather_gemm_hv": (_i, [_p, _i64, _i, _i64, _p, _p, _i, _i64, _p, _i, _p, _i, _p, _p, _f, _p, _i64, _i, _p, _i64, _i, _p, _sz,
                                _p, _p, _p, _p]),
    "fsfb_gather_gemm_simt": (_i, [_p, _i64, _i, _i64, _p, _i, _i64, _p, _i, _p, _i, _p, _p, _f, _p, _i64, _i, _p, _i64, _p]),
    "fsfb_conv_rulebook": (_i, [_p, _i64, _p, _p, _p, _p, _p, _p, _i, _p, _p]),
    "fsfb_conv_out_index": (_i, [_p, _i64, _p, _p, _p, _p, _p, _p, _sz, _p, _i64, _p, _p, _p]),
    "fsfb_ccl_workspace_bytes": (_i, [_i64, _psz]),
    "fsfb_connected_components": (_i, [_p, _i64, _i64, _p, _f, _p, _p, _p, _sz, _p]),
    "fsfb_vfe_decorate": (_i, [_p, _i64, _i, _i64, _p, _i, _p, _p, _p, _p, _i, _i, _p, _p]),
    "fsfb_sir_input": (_i, [_p, _i64, _i, _i64, _p, _p, _i64, _p, _i64, _p]),
    "fsfb_div_cols": (_i, [_p, _i64, _i, _i64, _p, _p, _i64, _p]),
    "fsfb_add_inplace": (_i, [_p, _i64, _i, _i64, _p, _i64, _p]),
    "fsfb_reduce_channel": (_i, [_p, _i64, _i, _i64, _i, _p, _p]),
    "fsfb_neck_points": (_i, [_p, _i64, _i64, _p, _i, _p, _i64, _i, _p, _i, _p, _p, _f, _p, _i64, _p, _p, _p]),
    "fsfb_vote_decode": (_i, [_p, _i64, _p, _p]),
    "fsfb_compact_workspace_bytes": (_i, [_i64, _psz]),
    "fsfb_compact_indices": (_i, [_p, _i64, _p, _p, _p, _sz, _p]),
    "fsfb_group_sample": (_i, [_p, _i64, _i, _p, _i64, _p, _p, _p, _i, _p, _p, _p, _p]),
    "fsfb_gather_overlap": (_i, [_p, _p, _i64, _p, _p]),
    "fsfb_frustum_expand": (_i, [_p, _i64, _p, _i, _p, _i, _i, _i, _i, _p, _i64, _p, _p, _p, _p, _p, _p, _p, _p]),
    "fsfb_weighted_xyz": (_i, [_p, _i64, _p, _p, _i64, _p, _p]),
    "fsfb_cluster_delta": (_i, [_p, _i64, _p, _i64, _p, _i64, _i, _p, _p, _p, _p]),
    "fsfb_encode_preds_2d": (_i, [_p, _i, _i, _p, _i, _i, _i64, _f, _f, _i, _p, _p, _p]),
    "fsfb_threshold_mask": (_i, [_p, _i64, _i64, _i, _f, _p, _p]),
    "fsfb_count_mask": (_i, [_p, _p, _i64, _i, _p, _p]),
    "fsfb_sir_gate_input": (_i, [_p, _i64, _i, _i64, _p, _i64, _i, _p, _i64, _f, _p, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _f, _i, _p, _i64, _p]),
    "fsfb_rulebook_order_workspace_bytes": (_i, [_i64, _i, _psz]),
    "fsfb_rulebook_row_order": (_i, [_p, _i, _i64, _p, _p, _sz, _p]),
    "fsfb_rownorm_act": (_i, [_p, _i64, _i, _i64, _p, _i, _p, _p, _f, _p, _i64, _i, _p, _i64, _p]),
}


class FsfbError(RuntimeError):
    pass


_lib = None


def load(build_if_missing: bool = True) -> C.CDLL:
    """Load (building first if the .so is absent and nvcc is available)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists() and build_if_missing and os.environ.get("FSFB_NO_AUTOBUILD") != "1":
        from . import build as _build

        _build.build()
    if not LIB_PATH.exists():
        raise FsfbError(
            f"{LIB_PATH} is missing: run `python -m fullysparsefusion_b200.build` "
            "(there is no CPU fallback for the hot path)"
        )
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here == header/library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != OK:
        msg = load().fsfb_last_error().decode("utf-8", "replace")
        kind = {ERR_BADARG: "bad argument", ERR_CAPACITY: "capacity", ERR_CUDA: "CUDA error"}.get(rc, str(rc))
        raise FsfbError(f"{what}: {kind}: {msg}")


def launch_count() -> int:
    return int(load().fsfb_launch_count())
