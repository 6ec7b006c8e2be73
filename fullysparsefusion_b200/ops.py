"""Torch-tensor front end of the C-ABI kernels.

PyTorch is used here for device memory (the caching allocator), streams and dtype bookkeeping
only; all arithmetic on the hot path happens in libfsf_b200.so.  Every function raises if a
tensor is not on a CUDA device — there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os
import weakref
from dataclasses import dataclass
from typing import Optional, Sequence, Tuple

import torch

from . import _capi
from ._capi import REDUCE_MAX, REDUCE_MEAN, REDUCE_SUM, check, load

_MODES = {"sum": REDUCE_SUM, "add": REDUCE_SUM, "mean": REDUCE_MEAN, "avg": REDUCE_MEAN, "max": REDUCE_MAX}


# Optional per-op CUDA-event timing (bench.py): set PROFILER to a list to collect
# (op family, start event, end event, algorithmic bytes, algorithmic flops) for the HBM-/tensor-bound ops.
PROFILER = None
PROFILE_ONLY = None  # optional set of op-family names: only these are timed (keeps event overhead out of a timed region)
DETAIL = False  # per-shape op names in the profile (tools only)


class _Prof:
    __slots__ = ("name", "nbytes", "flops", "a")

    def __init__(self, name: str, nbytes=0, flops=0):
        # nbytes / flops: ints, or (device scalar tensor, multiplier, constant) resolved by the reader after the run
        self.name, self.nbytes, self.flops, self.a = name, nbytes, flops, None

    def __enter__(self):
        if PROFILER is not None and (PROFILE_ONLY is None or self.name.split("[")[0] in PROFILE_ONLY):
            self.a = torch.cuda.Event(enable_timing=True)
            self.a.record()
        return self

    def __exit__(self, *exc):
        if self.a is not None:
            b = torch.cuda.Event(enable_timing=True)
            b.record()
            PROFILER.append((self.name, self.a, b, self.nbytes, self.flops))
        return False


def _need_cuda(*ts: torch.Tensor) -> torch.device:
    dev = None
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise _capi.FsfbError(
                "fullysparsefusion_b200 ops run on CUDA tensors only (no CPU fallback); got a "
                f"{t.device} tensor"
            )
        dev = dev or t.device
    return dev


_STREAM_CACHE = {}


def _stream(dev: torch.device) -> C.c_void_p:
    """Raw handle of torch's current stream on `dev`.  torch.cuda.current_stream builds a Stream object per call
    (~8 us, 300 calls per frame), so the handle is cached per (device, stream id) through the cheap C getter."""
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    raw = torch._C._cuda_getCurrentRawStream(idx)
    h = _STREAM_CACHE.get(raw)
    if h is None:
        h = _STREAM_CACHE[raw] = C.c_void_p(raw)
    return h


def _ptr(t: Optional[torch.Tensor]) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


def _ws(nbytes: int, dev: torch.device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=dev)


def empty_rows(n: int, c: int, dev: torch.device) -> torch.Tensor:
    """[n, c] fp32 whose rows start 16-byte aligned: wide odd widths (131, 133, 175 ...) are carved out of a
    [n, round_up(c, 4)] allocation so every consumer can use 128-bit accesses."""
    if c % 4 == 0 or c < 64:
        return torch.empty((n, c), dtype=torch.float32, device=dev)
    return torch.empty((n, (c + 3) // 4 * 4), dtype=torch.float32, device=dev)[:, :c]


def _padded_base(t: torch.Tensor) -> Optional[torch.Tensor]:
    """The [n, round_up(c,4)] allocation `t` is a [:, :c] view of (see empty_rows), or None."""
    b = t._base
    if (b is not None and t.dim() == 2 and t.size(1) >= 64 and b.dim() == 2 and b.is_contiguous() and b.data_ptr() == t.data_ptr()
            and b.size(0) == t.size(0) and b.size(1) == (t.size(1) + 3) // 4 * 4 and b.size(1) != t.size(1)
            and t.stride(0) == b.size(1) and t.stride(1) == 1):
        return b
    return None


def _host_f32(vals: Sequence[float]):
    return (C.c_float * len(vals))(*[float(v) for v in vals])


def _host_i64(vals: Sequence[int]):
    return (C.c_int64 * len(vals))(*[int(v) for v in vals])


def _host_i32(vals: Sequence[int]):
    return (C.c_int32 * len(vals))(*[int(v) for v in vals])


# ------------------------------------------------------------------------------------------
# a1 voxelize
# ------------------------------------------------------------------------------------------
def grid_shape(point_cloud_range: Sequence[float], voxel_size: Sequence[float]) -> Tuple[int, int, int]:
    """(gx, gy, gz) = round((max - min) / voxel) as mmdet3d's Voxelization computes it."""
    r, v = point_cloud_range, voxel_size
    return tuple(int(round((r[i + 3] - r[i]) / v[i])) for i in range(3))


def voxelize(points: torch.Tensor, voxel_size: Sequence[float], point_cloud_range: Sequence[float],
             floor_mode: int = 0, grid: Optional[Sequence[int]] = None, order_xyz: bool = False,
             check_range: bool = True) -> torch.Tensor:
    """Dynamic voxelization → coors [N,3] int32 (z,y,x); out-of-range → -1.

    Mirrors mmdet3d.ops.Voxelization(max_num_points=-1)(points)
    (reference call: projects/mmdet3d_plugin/models/detectors/single_stage_fsd.py:217-219).
    floor_mode=1 reproduces torch.div(..., rounding_mode='floor') (single_stage_fsd.py:270,591).
    """
    dev = _need_cuda(points)
    assert points.dim() == 2 and points.size(1) >= 3 and points.dtype == torch.float32
    if points.stride(1) != 1:
        points = points.contiguous()
    n = points.size(0)
    g = tuple(grid) if grid is not None else grid_shape(point_cloud_range, voxel_size)
    coors = torch.empty((n, 3), dtype=torch.int32, device=dev)
    rc = load().fsfb_voxelize(_ptr(points), n, points.stride(0) if n else 3, _host_f32(point_cloud_range[:3]),
                              _host_f32(voxel_size), _host_i32(g), int(floor_mode), int(order_xyz), int(check_range),
                              _ptr(coors), _stream(dev))
    check(rc, "fsfb_voxelize")
    return coors


# ------------------------------------------------------------------------------------------
# a2 row ranking (torch.unique(dim=0))
# ------------------------------------------------------------------------------------------
MAX_CELLS = (1 << 32) - 2


def rows_minmax(rows: torch.Tensor) -> Tuple[list, list]:
    """Per-column (min, max) of an integer [N,D] tensor; one device→host sync."""
    dev = _need_cuda(rows)
    n, d = rows.shape
    out = torch.empty(2 * d, dtype=torch.int64, device=dev)
    rc = load().fsfb_rows_minmax(_ptr(rows), int(rows.dtype == torch.int64), n, d, _ptr(out), _stream(dev))
    check(rc, "fsfb_rows_minmax")
    h = out.tolist()
    return h[:d], h[d:]


@dataclass
class VoxelIndex:
    """The ranked bitmap of an active-site set: coordinate → row by bit test + popcount prefix.
    `ws` is the workspace fsfb_rank_rows / fsfb_conv_out_index filled (bitmap at offset 0)."""
    ws: torch.Tensor
    lo: Tuple[int, ...]
    ext: Tuple[int, ...]
    m: int


def unique_rows(rows: torch.Tensor, lo: Optional[Sequence[int]] = None, ext: Optional[Sequence[int]] = None,
                return_counts: bool = False, return_unique: bool = True, inv_dtype: torch.dtype = torch.int64,
                return_index: bool = False):
    """torch.unique(rows, dim=0, return_inverse=True[, return_counts=True]) for bounded integer rows.

    Returns (unique_rows [M,D], inverse [N], counts [M] | None).  `lo`/`ext` give the per-column
    bounds when the caller knows them (a voxel grid); otherwise one extra min/max pass runs.
    Reference call sites: sst_ops.py:156,165; sir.py:68; single_stage_fsd.py:595.
    """
    dev = _need_cuda(rows)
    assert rows.dim() == 2 and rows.dtype in (torch.int64, torch.int32), (rows.shape, rows.dtype)
    rows = rows.contiguous()
    n, d = rows.shape
    if n == 0:   # torch.unique survives empty input: empty unique / inverse / counts (and an index over zero rows)
        res = (rows.new_empty((0, d)) if return_unique else None, torch.empty(0, dtype=inv_dtype, device=dev),
               torch.empty(0, dtype=torch.int32 if inv_dtype == torch.int32 else torch.int64, device=dev) if return_counts else None)
        if return_index:
            lo0 = tuple(int(v) for v in lo) if lo is not None else (0,) * d
            ext0 = tuple(int(v) for v in ext) if ext is not None else (1,) * d
            return res + (VoxelIndex(_ws(0, dev), lo0, ext0, 0),)
        return res
    if lo is None or ext is None:
        mn, mx = rows_minmax(rows)
        lo = mn
        ext = [b - a + 1 for a, b in zip(mn, mx)]
    cells = 1
    for e in ext:
        cells *= int(e)
    if cells > MAX_CELLS:
        raise _capi.FsfbError(
            f"unique_rows: key space of {cells} cells exceeds the bitmap ranker's 2^32-2 limit")
    lib = load()
    need = C.c_size_t(0)
    check(lib.fsfb_rank_workspace_bytes(n, cells, C.byref(need)), "fsfb_rank_workspace_bytes")
    ws = _ws(need.value, dev)
    inv32 = torch.empty(n, dtype=torch.int32, device=dev) if inv_dtype == torch.int32 else None
    inv64 = torch.empty(n, dtype=torch.int64, device=dev) if inv_dtype == torch.int64 else None
    uniq = torch.empty((n, d), dtype=rows.dtype, device=dev) if return_unique else None
    counts = torch.empty(n, dtype=torch.int32, device=dev) if return_counts else None
    meta = torch.empty(2, dtype=torch.int32, device=dev)  # [num_unique, status]
    with _Prof("rank_rows", n * d * rows.element_size() + n * (8 if inv64 is not None else 4)):
        rc = lib.fsfb_rank_rows(_ptr(rows), int(rows.dtype == torch.int64), n, d, _host_i64(lo), _host_i64(ext),
                                _ptr(ws), ws.numel(), _ptr(inv32), _ptr(inv64), _ptr(uniq), n, _ptr(counts),
                                C.c_void_p(meta.data_ptr()), C.c_void_p(meta.data_ptr() + 4), _stream(dev))
    check(rc, "fsfb_rank_rows")
    m, status = meta.tolist()  # the one sync torch.unique also pays (output size)
    if status & 1:
        raise _capi.FsfbError("unique_rows: a row lies outside the given lo/ext bounds")
    inv = inv64 if inv64 is not None else inv32
    if counts is not None:
        counts = counts[:m] if inv_dtype == torch.int32 else counts[:m].long()
    res = (uniq[:m] if uniq is not None else None, inv, counts)
    if return_index:
        return res + (VoxelIndex(ws, tuple(int(v) for v in lo), tuple(int(v) for v in ext), m),)
    return res


# ------------------------------------------------------------------------------------------
# segment CSR + reductions (torch_scatter)
# ------------------------------------------------------------------------------------------
@dataclass
class SegmentCSR:
    """The "scatter rulebook": rows grouped by segment id, stable inside a segment."""
    offsets: torch.Tensor  # [m+1] int32
    perm: torch.Tensor     # [n] int32 source row of each sorted position
    seg: torch.Tensor      # [n] int32 segment of each sorted position
    n: int
    m: int
    dense: bool = False  # every segment has at least one row (ids come from a ranking): skips the empty-segment pass


def build_csr(index: torch.Tensor, m: int) -> SegmentCSR:
    dev = _need_cuda(index)
    assert index.dim() == 1 and index.dtype in (torch.int64, torch.int32)
    index = index.contiguous()
    n = index.numel()
    if n == 0:   # nothing to sort: every segment is empty
        z = torch.zeros(0, dtype=torch.int32, device=dev)
        return SegmentCSR(torch.zeros(m + 1, dtype=torch.int32, device=dev), z, z, 0, m)
    lib = load()
    need = C.c_size_t(0)
    check(lib.fsfb_csr_workspace_bytes(n, m, C.byref(need)), "fsfb_csr_workspace_bytes")
    ws = _ws(need.value, dev)
    offsets = torch.empty(m + 1, dtype=torch.int32, device=dev)
    perm = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
    seg = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
    with _Prof("csr_build", index.element_size() * n + 8 * n + 4 * (m + 1)):
        rc = lib.fsfb_csr_build(_ptr(index), int(index.dtype == torch.int64), n, m, _ptr(offsets), _ptr(perm),
                                _ptr(seg), _ptr(ws), ws.numel(), _stream(dev))
    check(rc, "fsfb_csr_build")
    return SegmentCSR(offsets, perm[:n], seg[:n], n, m)


def segment_reduce(feat: torch.Tensor, csr: SegmentCSR, mode: str, return_argmax: bool = False):
    """out[s] = reduce(feat[rows of segment s]); mode in {'sum','mean','max'}.

    Equals torch_scatter.scatter(feat, index, 0, reduce=mode) / scatter_max (sst_ops.py:168,170);
    argmax ties resolve to the lowest source row; empty segments give 0 / argmax == N.
    """
    dev = _need_cuda(feat, csr.offsets)
    assert feat.dim() == 2 and feat.dtype == torch.float32 and feat.size(0) == csr.n
    base = _padded_base(feat)
    if base is not None and not return_argmax:   # reduce the padded rows with 128-bit accesses, drop the pad column
        return segment_reduce(base, csr, mode)[:, :feat.size(1)]
    if feat.stride(1) != 1 and feat.numel():
        feat = feat.contiguous()
    n, c = feat.shape
    mode_id = _MODES[mode]
    want_arg = return_argmax and mode_id == REDUCE_MAX
    mode_flags = mode_id | (0x100 if csr.dense else 0)
    out = torch.empty((csr.m, c), dtype=torch.float32, device=dev)
    arg = torch.empty((csr.m, c), dtype=torch.int64, device=dev) if want_arg else None
    if c == 0 or csr.m == 0:
        return (out, arg) if return_argmax else out
    lib = load()
    need = C.c_size_t(0)
    check(lib.fsfb_segment_reduce_workspace_bytes(n, c, int(want_arg), C.byref(need)),
          "fsfb_segment_reduce_workspace_bytes")
    ws = _ws(need.value, dev)
    with _Prof(f"segment_reduce[{mode},n={n},c={c}]", 4 * n * c + 8 * n + 4 * csr.m * c + (8 * csr.m * c if want_arg else 0)):
        rc = lib.fsfb_segment_reduce(_ptr(feat), n, c, feat.stride(0) if n else c, _ptr(csr.perm), _ptr(csr.seg),
                                     _ptr(csr.offsets), csr.m, mode_flags, _ptr(out), _ptr(arg), _ptr(ws), ws.numel(),
                                     _stream(dev))
    check(rc, "fsfb_segment_reduce")
    return (out, arg) if return_argmax else out


def gather_rows(src: torch.Tensor, idx: torch.Tensor, fill: float = 0.0, out: Optional[torch.Tensor] = None):
    """out[i] = src[idx[i]] (idx < 0 → fill).  voxel2point_neck.py:42-50, FSF.py:310-311."""
    dev = _need_cuda(src, idx)
    assert src.dim() == 2 and src.dtype == torch.float32
    assert idx.dim() == 1 and idx.dtype in (torch.int64, torch.int32)
    base = _padded_base(src)
    if base is not None and out is None:   # gather whole padded rows (128-bit), hand back the same padded view
        return gather_rows(base, idx, fill)[:, :src.size(1)]
    src = _rowmajor(src)
    idx = idx.contiguous()
    n, c = idx.numel(), src.size(1)
    if out is None:
        out = torch.empty((n, c), dtype=torch.float32, device=dev)
    assert out.size(0) == n and out.size(1) >= c and out.stride(1) == 1
    if n == 0 or c == 0:
        return out
    with _Prof(f"gather_rows[n={n},c={c}]", 2 * 4 * n * c + idx.element_size() * n):
        rc = load().fsfb_gather_rows(_ptr(src), src.size(0), c, src.stride(0) if src.size(0) else c, _ptr(idx),
                                     int(idx.dtype == torch.int64), n, float(fill), _ptr(out), out.stride(0), _stream(dev))
    check(rc, "fsfb_gather_rows")
    return out


def ingroup_indices(group: torch.Tensor, num_groups: Optional[int] = None) -> torch.Tensor:
    """Stable in-group index (ingroup_indices.forward, sst_ops.py:246-248)."""
    dev = _need_cuda(group)
    assert group.dim() == 1 and group.dtype == torch.int64
    group = group.contiguous()
    n = group.numel()
    out = torch.empty(n, dtype=torch.int64, device=dev)
    if n == 0:
        return out
    m = int(num_groups) if num_groups is not None else int(group.max().item()) + 1
    lib = load()
    need = C.c_size_t(0)
    check(lib.fsfb_ingroup_workspace_bytes(n, m, C.byref(need)), "fsfb_ingroup_workspace_bytes")
    ws = _ws(need.value, dev)
    rc = lib.fsfb_ingroup_indices(_ptr(group), n, m, _ptr(out), _ptr(ws), ws.numel(), _stream(dev))
    check(rc, "fsfb_ingroup_indices")
    return out


# ------------------------------------------------------------------------------------------
# a7+a8 projection + nearest sampling
# ------------------------------------------------------------------------------------------
def _mask_args(mask: torch.Tensor):
    assert mask.dim() == 4 and mask.dtype in (torch.uint8, torch.int32), (mask.shape, mask.dtype)
    return mask.contiguous(), int(mask.dtype == torch.int32)


def project_sample(xyz: torch.Tensor, lidar2img: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """FSF.points_in_mask (FSF.py:202-226): ids [N, cams, classes] int64."""
    dev = _need_cuda(xyz, lidar2img, mask)
    assert xyz.dim() == 2 and xyz.size(1) >= 3 and xyz.dtype == torch.float32
    if xyz.stride(1) != 1:
        xyz = xyz.contiguous()
    mask, is_i32 = _mask_args(mask)
    cams, classes, H, W = mask.shape
    l2i = lidar2img.to(torch.float32).contiguous()
    assert l2i.shape == (cams, 4, 4)
    n = xyz.size(0)
    out = torch.empty((n, cams, classes), dtype=torch.int64, device=dev)
    rc = load().fsfb_project_sample(_ptr(xyz), n, xyz.stride(0) if n else 3, _ptr(l2i), cams, _ptr(mask), is_i32,
                                    classes, H, W, _ptr(out), _stream(dev))
    check(rc, "fsfb_project_sample")
    return out


def project_sample_select(xyz: torch.Tensor, lidar2img: torch.Tensor, mask: torch.Tensor, want_overlap: bool = False,
                          anno: Optional[torch.Tensor] = None, anno_col: int = 4, want_ids: bool = True):
    """Fused contract: (ids_sel [N,classes] i32, cam_sel [N] u8, fg [N] u8[, overlap [N] u8][, scores [N,classes] f32]).
    scores = mask_anno[id-1][anno_col] of the camera-selected ids (FSF.img_cross_attn's MLP input on nuScenes)."""
    dev = _need_cuda(xyz, lidar2img, mask, anno)
    assert xyz.dim() == 2 and xyz.size(1) >= 3 and xyz.dtype == torch.float32
    if xyz.stride(1) != 1:
        xyz = xyz.contiguous()
    mask, is_i32 = _mask_args(mask)
    cams, classes, H, W = mask.shape
    l2i = lidar2img.to(torch.float32).contiguous()
    n = xyz.size(0)
    ids = torch.empty((n, classes), dtype=torch.int32, device=dev) if want_ids else None
    cam = torch.empty(n, dtype=torch.uint8, device=dev)
    fg = torch.empty(n, dtype=torch.uint8, device=dev)
    ov = torch.empty(n, dtype=torch.uint8, device=dev) if want_overlap else None
    scores = None
    a_rows = a_cols = 0
    if anno is not None:
        assert anno.dim() == 2 and anno.dtype == torch.float32
        anno = anno.contiguous()
        a_rows, a_cols = anno.shape
        scores = torch.empty((n, classes), dtype=torch.float32, device=dev)
    out_b = (4 * classes if want_ids else 0) + 2 + (1 if want_overlap else 0) + (4 * classes if anno is not None else 0)
    with _Prof("project_sample_select", n * (12 + cams * classes * mask.element_size() + out_b)):
        rc = load().fsfb_project_sample_select(_ptr(xyz), n, xyz.stride(0) if n else 3, _ptr(l2i), cams, _ptr(mask),
                                               is_i32, classes, H, W, _ptr(ids), _ptr(cam), _ptr(fg), _ptr(ov),
                                               _ptr(anno), a_rows, a_cols, int(anno_col), _ptr(scores), _stream(dev))
    check(rc, "fsfb_project_sample_select")
    out = (ids, cam, fg)
    if want_overlap:
        out += (ov,)
    if anno is not None:
        out += (scores,)
    return out


def project_sample_select_hwc(xyz: torch.Tensor, lidar2img: torch.Tensor, mask_hwc16: torch.Tensor, classes: int,
                              want_overlap: bool = False, anno: Optional[torch.Tensor] = None, anno_col: int = 4,
                              want_ids: bool = True):
    """EXPERIMENTAL twin of project_sample_select for class-interleaved planes: mask_hwc16 [cams, H, W, 16] u8 (byte k of a texel
    = id of class k; `loading.LoadMaskFromFiles(layout="hwc16")` produces it).  Same outputs."""
    dev = _need_cuda(xyz, lidar2img, mask_hwc16, anno)
    assert xyz.dim() == 2 and xyz.size(1) >= 3 and xyz.dtype == torch.float32
    if xyz.stride(1) != 1:
        xyz = xyz.contiguous()
    assert mask_hwc16.dim() == 4 and mask_hwc16.size(3) == 16 and mask_hwc16.dtype == torch.uint8 and mask_hwc16.is_contiguous()
    assert 1 <= classes <= 16
    cams, H, W, _ = mask_hwc16.shape
    l2i = lidar2img.to(torch.float32).contiguous()
    assert l2i.shape == (cams, 4, 4)
    n = xyz.size(0)
    ids = torch.empty((n, classes), dtype=torch.int32, device=dev) if want_ids else None
    cam = torch.empty(n, dtype=torch.uint8, device=dev)
    fg = torch.empty(n, dtype=torch.uint8, device=dev)
    ov = torch.empty(n, dtype=torch.uint8, device=dev) if want_overlap else None
    scores = None
    a_rows = a_cols = 0
    if anno is not None:
        assert anno.dim() == 2 and anno.dtype == torch.float32
        anno = anno.contiguous()
        a_rows, a_cols = anno.shape
        scores = torch.empty((n, classes), dtype=torch.float32, device=dev)
    out_b = (4 * classes if want_ids else 0) + 2 + (1 if want_overlap else 0) + (4 * classes if anno is not None else 0)
    with _Prof("project_sample_select_hwc", n * (12 + cams * classes + out_b)):
        rc = load().fsfb_project_sample_select_hwc(_ptr(xyz), n, xyz.stride(0) if n else 3, _ptr(l2i), cams, _ptr(mask_hwc16),
                                                   classes, H, W, _ptr(ids), _ptr(cam), _ptr(fg), _ptr(ov), _ptr(anno), a_rows,
                                                   a_cols, int(anno_col), _ptr(scores), _stream(dev))
    check(rc, "fsfb_project_sample_select_hwc")
    out = (ids, cam, fg)
    if want_overlap:
        out += (ov,)
    if anno is not None:
        out += (scores,)
    return out


# ------------------------------------------------------------------------------------------
# gather-GEMM (sparse conv / Linear) with fused epilogue
# ------------------------------------------------------------------------------------------
_NORMS = {None: _capi.NORM_NONE, "none": _capi.NORM_NONE, "ln": _capi.NORM_LAYERNORM, "affine": _capi.NORM_AFFINE}
_ACTS = {None: _capi.ACT_NONE, "none": _capi.ACT_NONE, "relu": _capi.ACT_RELU, "gelu": _capi.ACT_GELU}


@dataclass
class PackedWeight:
    """nn.Linear / sparse-conv weights [koff, cout, cin] in the tensor-core tile layout."""
    data: torch.Tensor  # uint8 blob written by fsfb_gemm_prepack
    koff: int
    cin: int
    cout: int
    raw: Optional[torch.Tensor] = None  # the fp32 [koff,cout,cin] tensor it was packed from


def gemm_prepack(w: torch.Tensor, keep_raw: bool = False) -> PackedWeight:
    """w: [cout,cin] (nn.Linear.weight) or [koff,cout,cin] f32 → PackedWeight."""
    dev = _need_cuda(w)
    assert w.dtype == torch.float32 and w.dim() in (2, 3)
    w3 = (w if w.dim() == 3 else w[None]).contiguous()
    koff, cout, cin = w3.shape
    lib = load()
    need = C.c_size_t(0)
    check(lib.fsfb_gemm_prepack_bytes(koff, cin, cout, C.byref(need)), "fsfb_gemm_prepack_bytes")
    data = torch.empty(need.value, dtype=torch.uint8, device=dev)
    check(lib.fsfb_gemm_prepack(_ptr(w3), koff, cin, cout, _ptr(data), _stream(dev)), "fsfb_gemm_prepack")
    return PackedWeight(data, koff, cin, cout, w3 if keep_raw else None)


_HOSTVEC = {}


def _host_vec(t: Optional[torch.Tensor]):
    """Host copy (ctypes float array) of a per-channel vector for fsfb_gather_gemm_hv.  Cached per tensor OBJECT (weak
    reference + in-place version counter), never per address: layer parameters are read back once, at the first call that
    uses them; a vector built on the fly is read back at every call (correct, just slower — keep such vectors on the module)."""
    if t is None:
        return None
    key = id(t)
    hit = _HOSTVEC.get(key)
    if hit is not None and hit[0]() is t and hit[1] == t._version:
        return hit[2]
    arr = (C.c_float * t.numel())(*t.detach().float().cpu().tolist())
    _HOSTVEC[key] = (weakref.ref(t, lambda _r, k=key: _HOSTVEC.pop(k, None)), t._version, arr)
    return arr


def _epilogue_args(cout, bias, norm, norm_w, norm_b, residual, act, dev):
    for t in (bias, norm_w, norm_b):
        assert t is None or (t.is_cuda and t.dtype == torch.float32 and t.numel() == cout and t.is_contiguous())
    if residual is not None:
        assert residual.dtype == torch.float32 and residual.dim() == 2 and residual.size(1) == cout and residual.stride(1) == 1
    return (_ptr(bias), _NORMS[norm], _ptr(norm_w), _ptr(norm_b))


_CHECK_STATUS = os.environ.get("FSFB_CHECK_STATUS", "0") != "0"
_SPLIT_ON = os.environ.get("FSFB_CONV_SPLIT", "1") != "0" and os.environ.get("FSFB_GEMM_SS", "1") != "0" \
    and os.environ.get("FSFB_GEMM_F16", "1") != "0"


def split_rows(a: torch.Tensor) -> torch.Tensor:
    """fp32 rows [n, c] (c % 32 == 0) → the fp16-split operand rows of fsfb_gather_gemm_split (uint8 [n, 4 c])."""
    dev = _need_cuda(a)
    assert a.dim() == 2 and a.dtype == torch.float32 and a.stride(1) == 1 and a.size(1) % 32 == 0
    out = torch.empty((a.size(0), 4 * a.size(1)), dtype=torch.uint8, device=dev)
    with _Prof("split_rows", 8 * a.size(0) * a.size(1)):
        check(load().fsfb_split_rows(_ptr(a), a.size(0), a.size(1), a.stride(0), _ptr(out), _stream(dev)), "fsfb_split_rows")
    return out


_SPLITS_AUTO = os.environ.get("FSFB_GEMM_AUTO_SPLITS", "1") != "0"
_LIN_KSPLIT = os.environ.get("FSFB_GEMM_LIN_KSPLIT", "1") != "0" and os.environ.get("FSFB_GEMM_LIN", "1") != "0"
_LIN_MIN_ROWS = int(os.environ.get("FSFB_GEMM_LIN_MIN_ROWS", "1024"))   # the same default as csrc/gemm_lin.cu


def linear_k_splits(rows: int, cin: int, cout: int) -> int:
    """K-chunk splits of a dense Linear layer whose (row tile, column tile) grid leaves most SMs idle while each tile walks a deep
    K (the 768 / 896 / 1024-wide refinement heads over a few thousand queries): the row-tile kernel (csrc/gemm_lin.cu) runs the K
    ranges as separate CTAs and `k_splitk_epilogue` sums the slabs and applies the epilogue (any LayerNorm width up to 1024)."""
    kc = (cin + 31) // 32
    if not _LIN_KSPLIT or (rows < _LIN_MIN_ROWS and cout > 256) or rows > 16384 or kc < 8 or cout > 1024:
        return 1
    tiles = ((rows + 127) // 128) * ((cout + 127) // 128)
    return max(1, min(kc // 4, 8, 296 // tiles))




def _pick_splits(rows: int, cpad: int, koff: int, kc: int) -> int:
    """Offset splits of a convolution on a level with few row tiles.  The persistent kernel runs ceil(units / 148) waves of
    (128-row, 128-column) units; splitting the offsets s ways makes the units s times shorter, so the last wave wastes less, at the
    price of the partial sums' round trip and one more launch.  Model: a unit costs ~0.62 us per (visited offset, K chunk) stage
    (1.2 k clk measured, ~60 % of the offsets visited on the deep levels); the split epilogue ~20 us + its traffic at 5 TB/s."""
    if koff < 6 or cpad > 1024 or koff * kc < 32:
        return 1
    units = ((rows + 127) // 128) * (cpad // 128)
    t_unit = koff * kc * 0.6 * 0.62e-6
    best_s, best_t = 1, -(-units // 148) * t_unit
    if not _SPLITS_AUTO:   # the round-1 rule: only when half the SMs would idle
        return max(1, min(koff // 3, 148 // units)) if units * 2 <= 148 else 1
    for s in range(2, min(koff // 3, 9) + 1):
        t = -(-units * s // 148) * t_unit / s + 20e-6 + 2.0 * rows * cpad * 4 * s / 5e12
        if t < 0.9 * best_t:
            best_s, best_t = s, t
    return best_s


def gather_gemm(a: torch.Tensor, w: PackedWeight, nbr: Optional[torch.Tensor] = None, rows: Optional[int] = None,
                bias=None, norm=None, norm_w=None, norm_b=None, eps: float = 1e-5, residual=None, act=None,
                out: Optional[torch.Tensor] = None, simt: bool = False, residual_post: bool = False,
                row_order: Optional[torch.Tensor] = None, splits: Optional[int] = None,
                nbr_ro: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[r] = act(norm(sum_k a[nbr[k][r]] @ w[k].T + bias) + residual)  (include/fsf_b200.h).

    splits: offset ranges run as independent work units (fsfb_gather_gemm_splitk); None picks it from the shape
    (only when rows/128 x cout/128 tiles would leave most SMs idle).

    nbr: int32 [koff, rows] neighbour table (< 0 = none) or None for a plain Linear over rows of `a`.
    simt=True runs the CUDA-core cross-check (tests only; needs w.raw).
    """
    dev = _need_cuda(a, w.data)
    assert a.dim() == 2 and a.dtype == torch.float32 and a.size(1) == w.cin, (a.shape, w.cin)
    if a.stride(1) != 1:
        a = a.contiguous()
    if nbr is not None:
        assert nbr.dtype == torch.int32 and nbr.dim() == 2 and nbr.size(0) == w.koff and nbr.is_contiguous()
        rows = nbr.size(1)
    else:
        assert w.koff == 1
        rows = a.size(0) if rows is None else rows
    if out is None:
        out = empty_rows(rows, w.cout, dev)
    assert out.size(0) == rows and out.size(1) == w.cout and out.stride(1) == 1
    if rows == 0:
        return out
    b, nid, nw, nb = _epilogue_args(w.cout, bias, norm, norm_w, norm_b, residual, act, dev)
    lib = load()
    if row_order is not None:
        assert row_order.dtype == torch.int32 and row_order.numel() == rows and row_order.is_contiguous()
    args = (_ptr(a), a.size(0), w.cin, a.stride(0), _ptr(nbr), w.koff, rows)
    tail = (w.cout, b, nid, nw, nb, float(eps), _ptr(residual), residual.stride(0) if residual is not None else 0,
            _ACTS[act] | (0x100 if residual_post else 0), _ptr(out), out.stride(0), _stream(dev))
    if simt:
        assert w.raw is not None, "gemm_prepack(..., keep_raw=True) needed for the SIMT cross-check"
        check(lib.fsfb_gather_gemm_simt(*args, _ptr(w.raw), *tail), "fsfb_gather_gemm_simt")
        return out
    if nbr is None:
        prof = _Prof(f"gather_gemm_linear_{rows >> 10}k_{w.cin}x{w.cout}" if DETAIL else "gather_gemm_linear", 4 * rows * (w.cin + w.cout) + 4 * w.cin * w.cout, 2 * rows * w.cin * w.cout)
    else:
        pairs = getattr(nbr, "_fsfb_pairs", None)
        if PROFILER is not None and pairs is None:   # profiling only: exact pair count of this rulebook, kept on device
            pairs = (nbr >= 0).sum()
            nbr._fsfb_pairs = pairs
        prof = _Prof(f"gather_gemm_conv_{rows >> 10}k_k{w.koff}_{w.cin}x{w.cout}" if DETAIL else "gather_gemm_conv", (pairs, 4 * w.cin, 4 * rows * w.cout + 4 * w.koff * w.cin * w.cout),
                     (pairs, 2 * w.cin * w.cout, 0))
    cpad = (w.cout + 127) // 128 * 128
    tileable = w.cout <= 128 or (((w.cout + 15) // 16 * 16) % 128 == 0 and norm != "layernorm" and norm != "ln")
    if splits is None:
        if nbr is None:
            splits = linear_k_splits(rows, w.cin, w.cout)
        else:
            splits = _pick_splits(rows, cpad, w.koff, (w.cin + 31) // 32) if tileable else 1
    # 27-offset convolutions: every input row is gathered by ~6 offsets, so the fp32 → fp16-split conversion is done once per
    # row up front (fsfb_split_rows) and the kernel's gather is pure data movement (include/fsf_b200.h)
    if (_SPLIT_ON and nbr is not None and w.koff > 1 and w.koff <= 27 and w.cin % 32 == 0 and tileable and a.data_ptr() % 16 == 0
            and a.stride(0) % 4 == 0 and a.size(0) > 0):
        a_s = split_rows(a)
        ws = torch.empty(splits * rows * cpad, dtype=torch.float32, device=dev) if splits > 1 else None
        hv = w.cout <= 128 and splits == 1
        if nbr_ro is not None and row_order is not None:   # the table permuted into the row order (permute_rulebook): contiguous tile tables
            assert nbr_ro.dtype == torch.int32 and nbr_ro.is_contiguous() and tuple(nbr_ro.shape) == (w.koff, (rows + 127) // 128 * 128)
            nbr = nbr_ro
            tail = tail[:8] + (tail[8] | 0x200,) + tail[9:]   # FSFB_NBR_ROW_ORDERED rides in `act`
        with prof:
            check(lib.fsfb_gather_gemm_split(_ptr(a_s), a.size(0), w.cin, _ptr(nbr), _ptr(row_order), w.koff, rows, _ptr(w.data), *tail[:-1],
                                             splits, _ptr(ws), ws.numel() * 4 if ws is not None else 0,
                                             _host_vec(bias) if hv else None, _host_vec(norm_w) if hv else None,
                                             _host_vec(norm_b) if hv else None, tail[-1]), "fsfb_gather_gemm_split")
        return out
    with prof:
        if splits > 1:
            ws = torch.empty(splits * rows * cpad, dtype=torch.float32, device=dev)
            check(lib.fsfb_gather_gemm_splitk(*args[:5], _ptr(row_order), *args[5:], _ptr(w.data), *tail[:-1], splits, _ptr(ws),
                                              ws.numel() * 4, tail[-1]), "fsfb_gather_gemm_splitk")
        elif w.cout <= 128 and (bias is not None or norm_w is not None or norm_b is not None):
            # per-channel vectors also as host copies: they ride in the kernel parameters (include/fsf_b200.h)
            check(lib.fsfb_gather_gemm_hv(*args[:5], _ptr(row_order), *args[5:], _ptr(w.data), *tail[:-1], 1, None, 0,
                                          _host_vec(bias), _host_vec(norm_w), _host_vec(norm_b), tail[-1]), "fsfb_gather_gemm_hv")
        else:
            check(lib.fsfb_gather_gemm(*args[:5], _ptr(row_order), *args[5:], _ptr(w.data), *tail), "fsfb_gather_gemm")
    return out


def rownorm_act(x: torch.Tensor, bias=None, norm=None, norm_w=None, norm_b=None, eps: float = 1e-5, residual=None,
                act=None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """y = act(norm(x + bias) + residual) row-wise (wide LayerNorms of the cluster heads)."""
    dev = _need_cuda(x)
    assert x.dim() == 2 and x.dtype == torch.float32 and x.stride(1) == 1
    rows, c = x.shape
    if out is None:
        out = torch.empty_like(x)
    b, nid, nw, nb = _epilogue_args(c, bias, norm, norm_w, norm_b, residual, act, dev)
    check(load().fsfb_rownorm_act(_ptr(x), rows, c, x.stride(0), b, nid, nw, nb, float(eps), _ptr(residual),
                                  residual.stride(0) if residual is not None else 0, _ACTS[act], _ptr(out),
                                  out.stride(0), _stream(dev)), "fsfb_rownorm_act")
    return out


# ------------------------------------------------------------------------------------------
# sparse-convolution rulebook
# ------------------------------------------------------------------------------------------
def _triple(v):
    return [int(v)] * 3 if isinstance(v, int) else [int(x) for x in v]


def conv_rulebook(out_coors: torch.Tensor, index: VoxelIndex, ksize=3, stride=1, pad=1, transposed: bool = False):
    """nbr [koff, m_out] int32 neighbour table (include/fsf_b200.h: fsfb_conv_rulebook).
    out_coors: int32 [m_out,4] (b,z,y,x); index: VoxelIndex of the INPUT site set."""
    dev = _need_cuda(out_coors, index.ws)
    assert out_coors.dtype == torch.int32 and out_coors.dim() == 2 and out_coors.size(1) == 4
    out_coors = out_coors.contiguous()
    k, s_, p = _triple(ksize), _triple(stride), _triple(pad)
    koff = k[0] * k[1] * k[2]
    m_out = out_coors.size(0)
    nbr = torch.empty((koff, m_out), dtype=torch.int32, device=dev)
    with _Prof("conv_rulebook", 16 * m_out + 4 * koff * m_out):
        rc = load().fsfb_conv_rulebook(_ptr(out_coors), m_out, _ptr(index.ws), _host_i64(index.lo), _host_i64(index.ext),
                                       _host_i32(k), _host_i32(s_), _host_i32(p), int(transposed), _ptr(nbr), _stream(dev))
    check(rc, "fsfb_conv_rulebook")
    return nbr


def conv_out_index(in_coors: torch.Tensor, out_shape_bzyx: Sequence[int], ksize=3, stride=2, pad=1):
    """Active output sites of a strided SparseConv3d: (out_coors int32 [m_out,4] ascending, VoxelIndex)."""
    dev = _need_cuda(in_coors)
    assert in_coors.dtype == torch.int32 and in_coors.dim() == 2 and in_coors.size(1) == 4
    in_coors = in_coors.contiguous()
    k, s_, p = _triple(ksize), _triple(stride), _triple(pad)
    ext = [int(v) for v in out_shape_bzyx]
    lo = [0, 0, 0, 0]
    cells = ext[0] * ext[1] * ext[2] * ext[3]
    lib = load()
    need = C.c_size_t(0)
    check(lib.fsfb_rank_workspace_bytes(0, cells, C.byref(need)), "fsfb_rank_workspace_bytes")
    ws = _ws(need.value, dev)
    m_in = in_coors.size(0)
    koff = k[0] * k[1] * k[2]
    cap = max(1, min(cells, m_in * koff))
    out = torch.empty((cap, 4), dtype=torch.int32, device=dev)
    meta = torch.empty(2, dtype=torch.int32, device=dev)
    rc = lib.fsfb_conv_out_index(_ptr(in_coors), m_in, _host_i64(lo), _host_i64(ext), _host_i32(k), _host_i32(s_),
                                 _host_i32(p), _ptr(ws), ws.numel(), _ptr(out), cap, C.c_void_p(meta.data_ptr()),
                                 C.c_void_p(meta.data_ptr() + 4), _stream(dev))
    check(rc, "fsfb_conv_out_index")
    m, status = meta.tolist()
    if status:
        raise _capi.FsfbError(f"conv_out_index: status {status}")
    return out[:m], VoxelIndex(ws, tuple(lo), tuple(ext), m)


# ------------------------------------------------------------------------------------------
# a14 connected components
# ------------------------------------------------------------------------------------------
def connected_components(points: torch.Tensor, batch_idx: Optional[torch.Tensor], dist: float,
                         return_count: bool = False):
    """Labels [m] int32 (single_stage_fsd.py:45-82 semantics; batch_idx=None → single-batch variant)."""
    dev = _need_cuda(points, batch_idx)
    assert points.dim() == 2 and points.size(1) >= 2 and points.dtype == torch.float32
    if points.stride(1) != 1:
        points = points.contiguous()
    m = points.size(0)
    if batch_idx is not None:
        assert batch_idx.numel() == m
        batch_idx = batch_idx.to(torch.int32).contiguous()
    labels = torch.empty(m, dtype=torch.int32, device=dev)
    count = torch.zeros(1, dtype=torch.int32, device=dev)
    lib = load()
    need = C.c_size_t(0)
    check(lib.fsfb_ccl_workspace_bytes(m, C.byref(need)), "fsfb_ccl_workspace_bytes")
    ws = _ws(need.value, dev)
    with _Prof("connected_components", 16 * m):
        rc = lib.fsfb_connected_components(_ptr(points), m, points.stride(0) if m else 3, _ptr(batch_idx), float(dist),
                                           _ptr(labels), _ptr(count), _ptr(ws), ws.numel(), _stream(dev))
    check(rc, "fsfb_connected_components")
    return (labels, count) if return_count else labels


# ------------------------------------------------------------------------------------------
# per-point fused passes
# ------------------------------------------------------------------------------------------
def _rowmajor(x: torch.Tensor) -> torch.Tensor:
    return x if x.stride(-1) == 1 else x.contiguous()


def vfe_decorate(features, coors, inv32, voxel_mean, voxel_size, point_cloud_range, with_cluster_center, with_voxel_center,
                 out: Optional[torch.Tensor] = None):
    dev = _need_cuda(features, coors)
    features = _rowmajor(features)
    n, cin = features.shape
    cout = cin + 3 * int(with_cluster_center) + 3 * int(with_voxel_center)
    if out is None:
        out = torch.empty((n, cout), dtype=torch.float32, device=dev)
    assert out.shape == (n, cout) and out.is_contiguous()
    coors = coors.contiguous()
    assert coors.dtype in (torch.int32, torch.int64) and coors.size(1) == 4
    if with_cluster_center:
        assert inv32.dtype == torch.int32 and voxel_mean.is_contiguous() and voxel_mean.size(1) == cin
    rc = load().fsfb_vfe_decorate(_ptr(features), n, cin, features.stride(0) if n else cin, _ptr(coors),
                                  int(coors.dtype == torch.int64), _ptr(inv32), _ptr(voxel_mean), _host_f32(voxel_size),
                                  _host_f32(point_cloud_range[:3]), int(with_cluster_center), int(with_voxel_center),
                                  _ptr(out), _stream(dev))
    check(rc, "fsfb_vfe_decorate")
    return out


def sir_input(features, xyz_normalizer, gate=None, out=None):
    dev = _need_cuda(features, gate)
    features = _rowmajor(features)
    n, c = features.shape
    if gate is not None:
        gate = _rowmajor(gate)
        assert gate.shape == (n, c)
    if out is None:
        out = torch.empty((n, c), dtype=torch.float32, device=dev)
    rc = load().fsfb_sir_input(_ptr(features), n, c, features.stride(0) if n else c, _host_f32(xyz_normalizer), _ptr(gate),
                               gate.stride(0) if gate is not None and n else c, _ptr(out), out.stride(0) if n else c,
                               _stream(dev))
    check(rc, "fsfb_sir_input")
    return out


_DIVISORS: dict = {}


def div_cols(x, divisors):
    """x[:, c] / divisors[c]  (f_cluster / rel_dist_scaler)."""
    dev = _need_cuda(x)
    x = _rowmajor(x)
    n, c = x.shape
    key = (dev, tuple(float(d) for d in divisors))
    if key not in _DIVISORS:
        _DIVISORS[key] = torch.tensor(key[1], dtype=torch.float32, device=dev)
    out = torch.empty((n, c), dtype=torch.float32, device=dev)
    rc = load().fsfb_div_cols(_ptr(x), n, c, x.stride(0) if n else c, _ptr(_DIVISORS[key]), _ptr(out), c, _stream(dev))
    check(rc, "fsfb_div_cols")
    return out


def add_(x, y):
    dev = _need_cuda(x, y)
    assert x.shape == y.shape and x.dim() == 2 and x.stride(1) == 1 and y.stride(1) == 1
    n, c = x.shape
    rc = load().fsfb_add_inplace(_ptr(x), n, c, x.stride(0) if n else c, _ptr(y), y.stride(0) if n else c, _stream(dev))
    check(rc, "fsfb_add_inplace")
    return x


def conv_wgrad(a: torch.Tensor, dy: torch.Tensor, nbr: Optional[torch.Tensor], koff: int) -> torch.Tensor:
    """Weight gradient of the gather-GEMM: dw[k][co][ci] = sum_r dy[r][co] * a[nbr[k][r]][ci]  → [koff, cout, cin] f32
    (include/fsf_b200.h; nbr None = Linear)."""
    dev = _need_cuda(a, dy)
    assert a.dim() == 2 and dy.dim() == 2 and a.dtype == torch.float32 and dy.dtype == torch.float32
    assert a.stride(1) == 1 and dy.stride(1) == 1
    if nbr is not None:
        assert nbr.dtype == torch.int32 and nbr.dim() == 2 and nbr.size(0) == koff and nbr.size(1) == dy.size(0) and nbr.is_contiguous()
    cin, cout, rows = a.size(1), dy.size(1), dy.size(0)
    dw = torch.empty((koff, cout, cin), dtype=torch.float32, device=dev)
    lib = load()
    need = C.c_size_t()
    check(lib.fsfb_conv_wgrad_workspace_bytes(rows, koff, cin, cout, C.byref(need)), "fsfb_conv_wgrad_workspace_bytes")
    ws = _ws(need.value, dev)
    with _Prof("conv_wgrad", 4 * (a.size(0) * cin + rows * cout + koff * cin * cout)):
        check(lib.fsfb_conv_wgrad(_ptr(a), a.size(0), cin, a.stride(0) if a.size(0) else cin, _ptr(dy), rows, cout,
                                  dy.stride(0) if rows else cout, _ptr(nbr), koff, _ptr(dw), _ptr(ws), ws.numel(), _stream(dev)),
              "fsfb_conv_wgrad")
    return dw


def reduce_channel(x, out_channels: int):
    dev = _need_cuda(x)
    assert x.dim() == 2 and x.stride(1) == 1
    n, cin = x.shape
    out = torch.empty((n, out_channels), dtype=torch.float32, device=dev)
    rc = load().fsfb_reduce_channel(_ptr(x), n, cin, x.stride(0) if n else cin, out_channels, _ptr(out), _stream(dev))
    check(rc, "fsfb_reduce_channel")
    return out


def compact_indices(mask: torch.Tensor) -> torch.Tensor:
    """Indices (int32, ascending) of the non-zero entries of a bool/uint8 mask; one D2H sync for the count."""
    dev = _need_cuda(mask)
    assert mask.dim() == 1 and mask.dtype in (torch.bool, torch.uint8)
    mask = mask.contiguous()
    n = mask.numel()
    idx = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
    count = torch.empty(1, dtype=torch.int32, device=dev)
    lib = load()
    need = C.c_size_t(0)
    check(lib.fsfb_compact_workspace_bytes(n, C.byref(need)), "fsfb_compact_workspace_bytes")
    ws = _ws(need.value, dev)
    check(lib.fsfb_compact_indices(_ptr(mask), n, _ptr(idx), _ptr(count), _ptr(ws), ws.numel(), _stream(dev)),
          "fsfb_compact_indices")
    return idx[: int(count.item())]


def voxel2point_neck(points, pts_coors, voxel_feats, voxel2point_inds, voxel_size, point_cloud_range, voxel_padding=-1.0):
    """Voxel2PointScatterNeck.forward → (results [N_kept, C+3], pts_mask [N] bool)."""
    dev = _need_cuda(points, pts_coors, voxel_feats, voxel2point_inds)
    points = _rowmajor(points)
    pts_coors = pts_coors.contiguous()
    voxel_feats = voxel_feats.contiguous()
    inv = voxel2point_inds.contiguous()
    n, (m, c) = points.size(0), voxel_feats.shape
    out = empty_rows(n, c + 3, dev)
    mask = torch.empty(n, dtype=torch.uint8, device=dev)
    dropped = torch.empty(1, dtype=torch.int32, device=dev)
    rc = load().fsfb_neck_points(_ptr(points), n, points.stride(0) if n else 3, _ptr(pts_coors),
                                 int(pts_coors.dtype == torch.int64), _ptr(voxel_feats), m, c, _ptr(inv),
                                 int(inv.dtype == torch.int64), _host_f32(voxel_size), _host_f32(point_cloud_range[:3]),
                                 float(voxel_padding), _ptr(out), out.stride(0), _ptr(mask), _ptr(dropped), _stream(dev))
    check(rc, "fsfb_neck_points")
    if int(dropped.item()):  # padded voxels present: compact exactly as the boolean indexing at :50-52 does
        keep = compact_indices(mask)
        out = gather_rows(out, keep)
    return out, mask.bool()


def vote_decode(preds: torch.Tensor) -> torch.Tensor:
    dev = _need_cuda(preds)
    preds = preds.contiguous()
    out = torch.empty_like(preds)
    check(load().fsfb_vote_decode(_ptr(preds), preds.numel(), _ptr(out), _stream(dev)), "fsfb_vote_decode")
    return out


# ------------------------------------------------------------------------------------------
# query-generation helpers
# ------------------------------------------------------------------------------------------
def group_sample(logits: torch.Tensor, groups: Optional[Sequence[Sequence[int]]] = None, xyz=None, offsets=None,
                 want_fg_weight: bool = False):
    """softmax-derived quantities of one [n, C+1] logit tensor (include/fsf_b200.h: fsfb_group_sample).
    Returns (fg_weight [n] | None, group_score [n,G] | None, group_center [n,G,3] | None)."""
    dev = _need_cuda(logits, xyz, offsets)
    logits = logits.contiguous()
    n, c1 = logits.shape
    groups = [list(g) for g in (groups or [])]
    G = len(groups)
    fgw = torch.empty(n, dtype=torch.float32, device=dev) if want_fg_weight else None
    score = torch.empty((n, G), dtype=torch.float32, device=dev) if G else None
    center = None
    if G and xyz is not None:
        xyz = _rowmajor(xyz)
        offsets = offsets.contiguous()
        assert offsets.numel() == n * c1 * 3
        center = torch.empty((n, G, 3), dtype=torch.float32, device=dev)
    flat = [c for g in groups for c in g]
    rc = load().fsfb_group_sample(_ptr(logits), n, c1, _ptr(xyz), xyz.stride(0) if xyz is not None and n else 3,
                                  _ptr(offsets), _host_i32([len(g) for g in groups]) if G else None,
                                  _host_i32(flat) if G else None, G, _ptr(fgw), _ptr(score), _ptr(center), _stream(dev))
    check(rc, "fsfb_group_sample")
    return fgw, score, center


def frustum_rows(xyz_noaug, lidar2img, mask, fg: torch.Tensor, overlap: torch.Tensor, batch_idx=None):
    """extract_fg_pts + double_overlap_pts + get_sir_coors (FSF.py:260-308, 357-365).
    Returns (rows_point int32 [R] — source point of every output row, sir_coors int32 [R,3] = (batch, 0, obj id),
    n_fg).  R = 0 when no point hits a mask (the caller fakes one object, FSF.py:407-414)."""
    dev = _need_cuda(xyz_noaug, lidar2img, mask, fg, overlap)
    idx_fg = compact_indices(fg if fg.dtype == torch.uint8 else fg.to(torch.uint8))
    n_fg = idx_fg.numel()
    if n_fg == 0:
        return torch.empty(0, dtype=torch.int32, device=dev), torch.empty((0, 3), dtype=torch.int32, device=dev), 0
    lib = load()
    ov32 = torch.empty(n_fg, dtype=torch.int32, device=dev)
    check(lib.fsfb_gather_overlap(_ptr(overlap), _ptr(idx_fg), n_fg, _ptr(ov32), _stream(dev)), "fsfb_gather_overlap")
    csr = build_csr(ov32, 18)
    off = csr.offsets.tolist()  # 19 ints: the one sync that sizes the output (the reference syncs per overlap count)
    if off[18] - off[17] > 0:
        raise _capi.FsfbError(f"frustum_rows: {off[18] - off[17]} point(s) lie in more than 16 (camera, class) masks; the expansion "
                              "keeps at most 16 object ids per point (csrc/project.cu kMaxOverlap)")
    extra = sum((off[k + 1] - off[k]) * (k - 1) for k in range(2, 17))
    rows = n_fg + extra
    xyz_noaug = _rowmajor(xyz_noaug)
    mask, is_i32 = _mask_args(mask)
    cams, classes, H, W = mask.shape
    l2i = lidar2img.to(torch.float32).contiguous()
    # zero-filled: should the kernel ever skip a row (status bit 4: re-sampled ids disagree with the overlap counts given) the
    # row points at point 0 / object 0 instead of at uninitialised memory
    rows_point = torch.zeros(rows, dtype=torch.int32, device=dev)
    sir_coors = torch.zeros((rows, 3), dtype=torch.int32, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    if batch_idx is not None:
        batch_idx = batch_idx.to(torch.int32).contiguous()
    rc = lib.fsfb_frustum_expand(_ptr(xyz_noaug), xyz_noaug.stride(0), _ptr(l2i), cams, _ptr(mask), is_i32, classes, H, W,
                                 _ptr(idx_fg), n_fg, _ptr(csr.perm), _ptr(csr.seg), _ptr(csr.offsets), _ptr(batch_idx),
                                 _ptr(rows_point), _ptr(sir_coors), _ptr(status), _stream(dev))
    check(rc, "fsfb_frustum_expand")
    if _CHECK_STATUS and int(status.item()) & 4:   # FSFB_CHECK_STATUS=1 (tests): costs a device sync per call
        raise _capi.FsfbError("frustum_rows: re-sampled object ids disagree with the overlap counts (status bit 4)")
    return rows_point, sir_coors, n_fg


def weighted_xyz(xyz, weight, rows: Optional[torch.Tensor] = None):
    dev = _need_cuda(xyz, weight)
    xyz = _rowmajor(xyz)
    n_rows = rows.numel() if rows is not None else xyz.size(0)
    out = torch.empty((n_rows, 4), dtype=torch.float32, device=dev)
    check(load().fsfb_weighted_xyz(_ptr(xyz), xyz.stride(0) if xyz.size(0) else 3, _ptr(weight.contiguous()), _ptr(rows),
                                   n_rows, _ptr(out), _stream(dev)), "fsfb_weighted_xyz")
    return out


def cluster_delta(xyz, mean, inv32, rows: Optional[torch.Tensor] = None):
    """(f_cluster [R,3], center [K,3]) from per-cluster means ([K,3] plain or [K,4] weighted sums)."""
    dev = _need_cuda(xyz, mean, inv32)
    xyz = _rowmajor(xyz)
    mean = mean.contiguous()
    k, mc = mean.shape
    n_rows = inv32.numel()
    center = torch.empty((k, 3), dtype=torch.float32, device=dev)
    f = torch.empty((n_rows, 3), dtype=torch.float32, device=dev)
    check(load().fsfb_cluster_delta(_ptr(xyz), xyz.stride(0) if xyz.size(0) else 3, _ptr(rows), n_rows, _ptr(mean), k, mc,
                                    _ptr(inv32), _ptr(center), _ptr(f), _stream(dev)), "fsfb_cluster_delta")
    return f, center


def encode_preds_2d(anno, obj_coors32, img_w, img_h, num_classes, coor_col: int = 2):
    """(preds_2d [K, 9], feat [K, 5+num_classes+1]) — get_single_cls_preds_2d + encode_preds_2d (FSF.py:449-504)."""
    dev = _need_cuda(anno, obj_coors32)
    anno = anno.contiguous()
    obj_coors32 = obj_coors32.contiguous()
    assert obj_coors32.dtype == torch.int32
    k = obj_coors32.size(0)
    preds = torch.empty((k, anno.size(1)), dtype=torch.float32, device=dev)
    feat = torch.empty((k, 5 + num_classes + 1), dtype=torch.float32, device=dev)
    check(load().fsfb_encode_preds_2d(_ptr(anno), anno.size(0), anno.size(1), _ptr(obj_coors32), obj_coors32.size(1), coor_col,
                                      k, float(img_w), float(img_h), num_classes, _ptr(preds), _ptr(feat), _stream(dev)),
          "fsfb_encode_preds_2d")
    return preds, feat


def gather_int_rows(src: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """Row gather of an int32 [m, c] tensor (a pure copy: routed through the fp32 row-gather kernel bit-for-bit)."""
    assert src.dtype == torch.int32 and src.dim() == 2
    return gather_rows(src.contiguous().view(torch.float32), idx).view(torch.int32)


def threshold_mask(x: torch.Tensor, col: int, thr: float) -> torch.Tensor:
    dev = _need_cuda(x)
    assert x.dim() == 2 and x.stride(1) == 1 and x.dtype == torch.float32
    n = x.size(0)
    mask = torch.empty(n, dtype=torch.uint8, device=dev)
    check(load().fsfb_threshold_mask(_ptr(x), n, x.stride(0) if n else x.size(1), col, float(thr), _ptr(mask), _stream(dev)),
          "fsfb_threshold_mask")
    return mask


def count_mask(counts32: torch.Tensor, inv32: torch.Tensor, min_count: int) -> torch.Tensor:
    dev = _need_cuda(counts32, inv32)
    assert counts32.dtype == torch.int32 and inv32.dtype == torch.int32
    n = inv32.numel()
    mask = torch.empty(n, dtype=torch.uint8, device=dev)
    check(load().fsfb_count_mask(_ptr(counts32.contiguous()), _ptr(inv32.contiguous()), n, int(min_count), _ptr(mask),
                                 _stream(dev)), "fsfb_count_mask")
    return mask


def sir_gate_input(features, f_cluster, rel_dist_scaler, xyz_normalizer, layers, eps: float, act: str, out=None,
                   features_b=None):
    """Fused SIRLayer input: cat(xyz/normalizer, feats) * rel_mlp(f_cluster / scaler)  (include/fsf_b200.h).
    layers: [(W1, ln_w1, ln_b1), (W2, ln_w2, ln_b2), (W3, ln_w3, ln_b3)] contiguous fp32 CUDA tensors."""
    dev = _need_cuda(features, f_cluster)
    features = _rowmajor(features)
    f_cluster = _rowmajor(f_cluster)
    n, c_a = features.shape
    c = c_a
    if features_b is not None:   # input row = cat(features, features_b) without materialising it
        features_b = _rowmajor(features_b)
        assert features_b.size(0) == n
        c = c_a + features_b.size(1)
    (w1, g1, b1), (w2, g2, b2), (w3, g3, b3) = layers
    h1, h2 = w1.size(0), w2.size(0)
    assert w1.shape == (h1, 3) and w2.shape == (h2, h1) and w3.shape == (c, h2)
    if out is None:
        out = torch.empty((n, c), dtype=torch.float32, device=dev)
    with _Prof("sir_gate_input", 4 * n * (2 * c + 3)):
        rc = load().fsfb_sir_gate_input(_ptr(features), n, c, features.stride(0) if n else c, _ptr(features_b),
                                        features_b.stride(0) if features_b is not None and n else 0, c_a, _ptr(f_cluster),
                                        f_cluster.stride(0) if n else 3, float(rel_dist_scaler), _host_f32(xyz_normalizer),
                                        h1, h2, _ptr(w1), _ptr(g1), _ptr(b1), _ptr(w2), _ptr(g2), _ptr(b2), _ptr(w3), _ptr(g3),
                                        _ptr(b3), float(eps), _ACTS[act], _ptr(out), out.stride(0) if n else c, _stream(dev))
    check(rc, "fsfb_sir_gate_input")
    return out


def permute_rulebook(nbr: torch.Tensor, order: torch.Tensor) -> torch.Tensor:
    """nbr [koff, rows] and a row order → the table in that order, padded to whole 128-row tiles with -1
    ([koff, round_up(rows, 128)] i32): what FSFB_NBR_ROW_ORDERED expects (include/fsf_b200.h).  Once per rulebook."""
    dev = _need_cuda(nbr, order)
    assert nbr.dtype == torch.int32 and nbr.dim() == 2 and nbr.is_contiguous() and order.dtype == torch.int32 and order.numel() == nbr.size(1)
    koff, rows = nbr.shape
    out = torch.empty((koff, (rows + 127) // 128 * 128), dtype=torch.int32, device=dev)
    check(load().fsfb_permute_rulebook(_ptr(nbr), koff, rows, _ptr(order), _ptr(out), _stream(dev)), "fsfb_permute_rulebook")
    return out


def rulebook_row_order(nbr: torch.Tensor) -> torch.Tensor:
    """Permutation of a rulebook's output rows sorted by their mask of present offsets (int32 [rows])."""
    dev = _need_cuda(nbr)
    assert nbr.dtype == torch.int32 and nbr.dim() == 2 and nbr.is_contiguous()
    koff, rows = nbr.shape
    order = torch.empty(rows, dtype=torch.int32, device=dev)
    lib = load()
    need = C.c_size_t(0)
    check(lib.fsfb_rulebook_order_workspace_bytes(rows, koff, C.byref(need)), "fsfb_rulebook_order_workspace_bytes")
    ws = _ws(need.value, dev)
    with _Prof("rulebook_row_order", 4 * koff * rows + 4 * rows):
        check(lib.fsfb_rulebook_row_order(_ptr(nbr), koff, rows, _ptr(order), _ptr(ws), ws.numel(), _stream(dev)),
              "fsfb_rulebook_row_order")
    return order


# ------------------------------------------------------------------------------------------
# f1 dynamic point pooling (query refinement)
# ------------------------------------------------------------------------------------------
def dynamic_point_pool(rois: torch.Tensor, pts: torch.Tensor, extra_wlh: Sequence[float], max_inbox_point: int,
                       out_pts_idx: torch.Tensor, out_roi_idx: torch.Tensor, out_pts_feats: torch.Tensor) -> torch.Tensor:
    """dynamic_point_pool_ext.forward (projects/mmdet3d_plugin/ops/dynamic_point_pool_op.py:27-32): fills the caller's
    prefilled buffers in canonical (roi, point) order and returns the device count [1] i32 of rows written."""
    dev = _need_cuda(rois, pts, out_pts_idx, out_roi_idx, out_pts_feats)
    assert rois.dim() == 2 and rois.size(1) == 7 and rois.dtype == torch.float32
    assert pts.dim() == 2 and pts.size(1) >= 3 and pts.dtype == torch.float32
    assert out_pts_idx.dtype == torch.int64 and out_roi_idx.dtype == torch.int64 and out_pts_feats.dtype == torch.float32
    assert out_pts_idx.is_contiguous() and out_roi_idx.is_contiguous() and out_pts_feats.is_contiguous()
    cap = out_pts_idx.numel()
    assert out_roi_idx.numel() == cap and tuple(out_pts_feats.shape) == (cap, 13)
    rois = rois.contiguous()
    if pts.stride(1) != 1:
        pts = pts.contiguous()
    lib = load()
    need = C.c_size_t()
    check(lib.fsfb_dynamic_point_pool_workspace_bytes(rois.size(0), int(max_inbox_point), C.byref(need)),
          "fsfb_dynamic_point_pool_workspace_bytes")
    ws = _ws(need.value, dev)
    num = torch.empty(1, dtype=torch.int32, device=dev)
    with _Prof("dynamic_point_pool", 12 * pts.size(0) + 28 * rois.size(0) + 68 * cap):
        check(lib.fsfb_dynamic_point_pool(_ptr(rois), rois.size(0), _ptr(pts), pts.size(0), pts.stride(0) if pts.size(0) else 3,
                                          _host_f32(extra_wlh), int(max_inbox_point), cap, _ptr(out_pts_idx), _ptr(out_roi_idx),
                                          _ptr(out_pts_feats), _ptr(num), _ptr(ws), ws.numel(), _stream(dev)),
              "fsfb_dynamic_point_pool")
    return num


# ------------------------------------------------------------------------------------------
# a13/a14 all class groups at once (csrc/group_cluster.cu)
# ------------------------------------------------------------------------------------------
def group_cluster(score: torch.Tensor, centers: torch.Tensor, thresholds: Sequence[float], voxel_sizes: Sequence[Sequence[float]],
                  range_min: Sequence[float], dists: Sequence[float], min_points: int):
    """group_sample's per-group selection + ClusterAssigner.forward_single_class for every class group in one pass
    (single_stage_fsd.py:822-842, 936-982).  score [n,G], centers [n,G,3] →
        rows [P] i32 (row of the per-voxel tensors), grp [P] i32 (class group), clu [P] i32 (cluster id inside the group),
        center_preds [P,3] f32 — ordered (group, row), exactly the concatenation the reference's loop builds."""
    dev = _need_cuda(score, centers)
    n, G = score.shape
    assert score.dtype == torch.float32 and score.stride(1) == 1 and centers.is_contiguous() and centers.numel() == n * G * 3
    lib, st = load(), _stream(dev)
    flags = torch.empty(G * n, dtype=torch.uint8, device=dev)
    small = torch.empty(3 * 8, dtype=torch.int32, device=dev)   # counts | kept | base
    check(lib.fsfb_group_flags(_ptr(score), n, score.stride(0) if n else G, G, _host_f32(thresholds), _ptr(flags), _ptr(small), st),
          "fsfb_group_flags")
    flat = compact_indices(flags)                                                        # list length 1
    t = flat.numel()
    grp, vox, cidx = (torch.empty(t, dtype=torch.int32, device=dev) for _ in range(3))
    check(lib.fsfb_group_split(_ptr(flat), t, n, G, _ptr(grp), _ptr(vox), _ptr(cidx), st), "fsfb_group_split")
    ctr = gather_rows(centers.view(n * G, 3), cidx)
    rows4 = torch.empty((t, 4), dtype=torch.int32, device=dev)
    check(lib.fsfb_group_voxelize(_ptr(ctr), t, _ptr(grp), _host_f32(range_min[:3]), _host_f32([v for vs in voxel_sizes for v in vs]),
                                  G, _ptr(rows4), st), "fsfb_group_voxelize")
    _, inv, cnt = unique_rows(rows4, return_counts=True, return_unique=False, inv_dtype=torch.int32)   # list length 2
    keep_flag = torch.empty(t, dtype=torch.uint8, device=dev)
    check(lib.fsfb_group_keep(_ptr(cnt), _ptr(inv), _ptr(grp), t, G, int(min_points), _ptr(keep_flag), _ptr(small[8:]), st),
          "fsfb_group_keep")
    keep = compact_indices(keep_flag)                                                    # list length 3
    p = keep.numel()
    ctr_k = gather_rows(ctr, keep)
    rows4_k = gather_int_rows(rows4, keep)
    sel = gather_int_rows(torch.stack([vox, grp], dim=1), keep)
    new_coors, inv_c, _ = unique_rows(rows4_k, inv_dtype=torch.int32)[:3]                # list length 4
    m = new_coors.size(0)
    csr = build_csr(inv_c, m)
    csr.dense = True
    sampled_centers = segment_reduce(ctr_k, csr, "mean")
    batch_c = new_coors[:, 0].contiguous()
    labels = torch.empty(m, dtype=torch.int32, device=dev)
    need = C.c_size_t(0)
    check(lib.fsfb_ccl_workspace_bytes(m, C.byref(need)), "fsfb_ccl_workspace_bytes")
    ws = _ws(need.value, dev)
    with _Prof("connected_components", 16 * m):
        check(lib.fsfb_connected_components_groups(_ptr(sampled_centers), m, sampled_centers.stride(0) if m else 3, _ptr(batch_c),
                                                   _host_f32(dists), G, _ptr(labels), None, _ptr(ws), ws.numel(), st),
              "fsfb_connected_components_groups")
    clu = torch.empty(p, dtype=torch.int32, device=dev)
    check(lib.fsfb_group_relabel(_ptr(labels), _ptr(batch_c), 1, m, G, _ptr(inv_c), p, _ptr(small[16:]), _ptr(clu), st),
          "fsfb_group_relabel")
    return sel[:, 0].contiguous(), sel[:, 1].contiguous(), clu, ctr_k


def decode_boxes(reg: torch.Tensor, base_points: torch.Tensor, batch: Optional[torch.Tensor] = None) -> torch.Tensor:
    """FSF.decode_stage_bboxes (FSF.py:1085-1095): rois [K, code] = (batch, xyz, dims, yaw[, vx, vy]) from regression rows."""
    dev = _need_cuda(reg, base_points, batch)
    reg = _rowmajor(reg)
    base_points = _rowmajor(base_points)
    k, code = reg.shape
    rois = torch.empty((k, code), dtype=torch.float32, device=dev)
    if batch is not None:
        assert batch.dtype == torch.int32 and batch.dim() == 1
    check(load().fsfb_decode_boxes(_ptr(reg), k, code, reg.stride(0) if k else code, _ptr(base_points),
                                   base_points.stride(0) if k else 3, _ptr(batch), batch.stride(0) if batch is not None and k else 1,
                                   _ptr(rois), _stream(dev)), "fsfb_decode_boxes")
    return rois


def multiclass_nms(boxes: torch.Tensor, logits: torch.Tensor, score_thr: float, nms_thr: float, max_num: int,
                   apply_sigmoid: bool = True):
    """box3d_multiclass_nms with rotated BEV IoU on sigmoid(logits) (frustum_cluster_head.py:595-698).
    boxes [K, >=7] (x,y,z,dx,dy,dz,yaw,...), logits [K, C] → (boxes [n, D], scores [n], labels [n] i64, source rows [n] i32)."""
    dev = _need_cuda(boxes, logits)
    boxes, logits = _rowmajor(boxes), _rowmajor(logits)
    k, nc = logits.shape
    D = boxes.size(1)
    lib, st = load(), _stream(dev)
    empty = (torch.empty((0, D), device=dev), torch.empty(0, device=dev), torch.empty(0, dtype=torch.int64, device=dev),
             torch.empty(0, dtype=torch.int32, device=dev))
    if k == 0:
        return empty
    scores = torch.empty((k, nc), dtype=torch.float32, device=dev)
    flags = torch.empty(nc * k, dtype=torch.uint8, device=dev)
    counts = torch.empty(nc, dtype=torch.int32, device=dev)
    check(lib.fsfb_nms_flags(_ptr(logits), k, nc, logits.stride(0), int(apply_sigmoid), float(score_thr), _ptr(scores), _ptr(flags),
                             _ptr(counts), st), "fsfb_nms_flags")
    flat = compact_indices(flags)
    T = flat.numel()
    if T == 0:
        return empty
    max_class = max(counts.tolist())
    need = C.c_size_t(0)
    check(lib.fsfb_nms_workspace_bytes(T, max_class, C.byref(need)), "fsfb_nms_workspace_bytes")
    ws = _ws(need.value, dev)
    keep = torch.empty(T, dtype=torch.uint8, device=dev)
    check(lib.fsfb_nms_suppress(_ptr(boxes), k, boxes.stride(0), _ptr(scores), nc, _ptr(flat), T, _ptr(counts), max_class, float(nms_thr),
                                _ptr(keep), _ptr(ws), ws.numel(), st), "fsfb_nms_suppress")
    kept_idx = compact_indices(keep)
    P = kept_idx.numel()
    n = min(P, int(max_num))
    out_boxes = torch.empty((n, D), dtype=torch.float32, device=dev)
    out_scores = torch.empty(n, dtype=torch.float32, device=dev)
    out_labels = torch.empty(n, dtype=torch.int64, device=dev)
    out_rows = torch.empty(n, dtype=torch.int32, device=dev)
    check(lib.fsfb_nms_emit(_ptr(boxes), boxes.stride(0), D, _ptr(kept_idx), P, T, max_class, int(max_num), _ptr(out_boxes),
                            _ptr(out_scores), _ptr(out_labels), _ptr(out_rows), _ptr(ws), ws.numel(), st), "fsfb_nms_emit")
    return out_boxes, out_scores, out_labels, out_rows
