// point_pool.cu — dynamic point pooling for the query-refinement stage.
//
// Reference call: dynamic_point_pool_ext.forward(rois, pts, extra_wlh, max_inbox_point, out_pts_idx, out_roi_idx,
// out_pts_feats)  (projects/mmdet3d_plugin/ops/dynamic_point_pool_op.py:27-32), driven per sample by
// DynamicPointROIExtractor.forward (models/roi_heads/roi_extractors/dynamic_point_roi_extractor.py:30-100) from
// FSF.query_feat_refine (models/detectors/FSF.py:1020-1024).  The extension's source is not vendored (modified mmdet3d
// fork); the arithmetic below restates the published FSD kernel and satisfies every invariant the extractor asserts
// in-tree (:84-92): feats[0:3] = point, offsets pair up to the box dims (l, w, h), |local| inside dims + extra.
//
//   roi   = (cx, cy, cz, w, l, h, rz), centre = gravity centre (FSF.decode_stage_bboxes, FSF.py:1085-1095)
//   local = Rz(-rz) (p - c);  inside the ENLARGED box: |lx| < (l+e0)/2, |ly| < (w+e1)/2, |lz| <= (h+e2)/2
//   feats = [x, y, z, lx, ly, lz, lx + l/2, ly + w/2, lz + h/2, l/2 - lx, w/2 - ly, h/2 - lz, in_margin]
//   in_margin = 1 when the point is outside the original box (inside only thanks to extra_wlh)
//
// Upstream appends (point, roi) hits with atomics, so its output order and which points survive the per-roi cap are
// run-dependent.  Canonical order here: roi-major, point index ascending; the per-roi cap keeps the lowest point
// indices, the global cap (the caller's buffer length, 50000 upstream) keeps the first entries of that order.
//
// B200 design: the K x N scan runs thread-per-roi over point slices (k_pp_scan: a count pass and a store pass, broadcast
// shared-memory reads, no atomics, canonical order by construction) and leaves <= max_inbox point ids per roi in a
// scratch table; a one-CTA scan turns counts into output offsets, and k_pp_write (warp per roi) writes ids and the 13
// features.
#include "common.cuh"

namespace fsfb {

constexpr int kPpWarps = 8;

struct PpBox {
  float cx, cy, cz, hl, hw, hh, el, ew, eh, cosa, sina, r2;
};

__device__ __forceinline__ PpBox pp_box(const float* __restrict__ roi, float e0, float e1, float e2) {
  PpBox b;
  b.cx = roi[0]; b.cy = roi[1]; b.cz = roi[2];
  const float w = roi[3], l = roi[4], h = roi[5], rz = roi[6];
  b.hl = l * 0.5f; b.hw = w * 0.5f; b.hh = h * 0.5f;
  b.el = (l + e0) * 0.5f; b.ew = (w + e1) * 0.5f; b.eh = (h + e2) * 0.5f;
  // mmdet3d 0.x lidar_to_local_coords (roiaware_pool3d / the fork's dynamic_point_pool): rotate the offset by rz + pi/2, l along
  // local x, w along local y — the rectangle is the axis-aligned (w along x, l along y) box turned CLOCKWISE by rz, the same
  // footprint nms.cu's rect_from_box gives the box (in-tree anchor: fsd_bbox_head_fsd.py:307-309 maps local → global with
  // rotation_3d_in_axis(local, roi_ry + pi/2))
  const float rot = __fadd_rn(rz, 1.57079632679489661923f);
  b.cosa = cosf(rot);
  b.sina = sinf(rot);
  b.r2 = (b.el * b.el + b.ew * b.ew) * 1.0001f + 1e-6f;   // circumscribed circle of the enlarged footprint (with rounding slack)
  return b;
}

__device__ __forceinline__ bool pp_local(const PpBox& b, float x, float y, float z, float& lx, float& ly, float& lz) {
  const float sx = __fsub_rn(x, b.cx), sy = __fsub_rn(y, b.cy);
  lz = __fsub_rn(z, b.cz);
  // a point outside the footprint's circumscribed circle cannot be inside the box: the brute-force K x N scan rejects almost every
  // pair here, before the rotation (the accepted pairs run the exact test below, unchanged)
  if (sx * sx + sy * sy > b.r2) {
    lx = ly = 0.f;
    return false;
  }
  // no FMA contraction: the oracle evaluates the same four products and two sums in fp32
  lx = __fadd_rn(__fmul_rn(sx, b.cosa), __fmul_rn(sy, -b.sina));
  ly = __fadd_rn(__fmul_rn(sx, b.sina), __fmul_rn(sy, b.cosa));
  return (fabsf(lz) <= b.eh) & (lx > -b.el) & (lx < b.el) & (ly > -b.ew) & (ly < b.ew);
}

// Thread = 4 rois, CTA = 256 rois x one slice of the points.  The slice streams through shared memory as (x, y, z, x^2 + y^2)
// tiles; every thread reads the same point (one broadcast 128-bit read per 128 rois) and rejects it per roi with
// |p|^2 - 2 p.c <= r^2 - |c|^2 + slack — two FMAs and a compare, deliberately loose: it may only over-select — and the rare
// survivors run pp_local, the exact test (which starts with the tight circle).  A thread meets its points in ascending order
// and slices are ascending ranges, so (slice, local order) is the canonical ascending point order with no sort and no atomics:
//   pass 1 (WRITE = false) counts the hits of every (slice, roi); k_pp_slice_prefix turns them into exclusive prefixes;
//   pass 2 (WRITE = true) re-scans and stores the ids at prefix + local rank, while below max_inbox.
constexpr int kPpThreads = 64;
constexpr int kPpPerThread = 4;
constexpr int kPpRois = kPpThreads * kPpPerThread;   // rois per CTA
constexpr int kPpTile = 1024;                        // points per shared-memory tile (16 KB)
constexpr int kPpMaxSlices = 256;

template <bool WRITE>
__global__ void __launch_bounds__(kPpThreads)
    k_pp_scan(const float* __restrict__ rois, int64_t k, const float* __restrict__ pts, int64_t n, int64_t pts_stride,
              float e0, float e1, float e2, int max_inbox, int64_t slice_len, int32_t* __restrict__ slice_counts,
              int32_t* __restrict__ scratch) {
  __shared__ float4 s_pts[kPpTile];
  const int slice = blockIdx.y;
  PpBox b[kPpPerThread];
  float m2x[kPpPerThread], m2y[kPpPerThread], lim[kPpPerThread];
  int64_t r[kPpPerThread];
  int before[kPpPerThread], cnt[kPpPerThread];
  bool on[kPpPerThread];
  bool any = false;
#pragma unroll
  for (int j = 0; j < kPpPerThread; ++j) {
    r[j] = (int64_t)blockIdx.x * kPpRois + j * kPpThreads + threadIdx.x;   // warp-contiguous rois per j: coalesced tables
    const bool have = r[j] < k;
    b[j] = pp_box(rois + (have ? r[j] : 0) * 7, e0, e1, e2);
    m2x[j] = -2.f * b[j].cx;
    m2y[j] = -2.f * b[j].cy;
    // slack: the two forms differ by rounding of terms of size |p|^2, |c|^2 (<= ~1e5 m^2 in a +-200 m scene: ~1e-2 absolute)
    const float cc = b[j].cx * b[j].cx + b[j].cy * b[j].cy;
    lim[j] = b[j].r2 - cc + 1e-6f * (cc + b[j].r2) + 0.05f;
    before[j] = (WRITE && have) ? slice_counts[(int64_t)slice * k + r[j]] : 0;
    cnt[j] = 0;
    on[j] = have && (!WRITE || before[j] < max_inbox);
    if (WRITE && on[j]) {   // nothing to store when this (slice, roi) had no hit in pass 1
      const int nxt = slice + 1 < (int)gridDim.y ? slice_counts[(int64_t)(slice + 1) * k + r[j]] : -1;
      if (nxt == before[j]) on[j] = false;
    }
    if (!on[j]) lim[j] = -INFINITY;
    any |= on[j];
  }
  if (WRITE && !__syncthreads_or(any)) return;
  const int64_t p0 = (int64_t)slice * slice_len, p1 = min(n, p0 + slice_len);
  for (int64_t t0 = p0; t0 < p1; t0 += kPpTile) {
    const int tn = (int)min((int64_t)kPpTile, p1 - t0);
    __syncthreads();
    for (int i = threadIdx.x; i < tn; i += kPpThreads) {
      const float* p = pts + (t0 + i) * pts_stride;
      const float x = __ldg(p), y = __ldg(p + 1);
      s_pts[i] = make_float4(x, y, __ldg(p + 2), x * x + y * y);
    }
    __syncthreads();
    if (!any) continue;
#pragma unroll 2
    for (int i = 0; i < tn; ++i) {
      const float4 q = s_pts[i];
#pragma unroll
      for (int j = 0; j < kPpPerThread; ++j) {
        if (fmaf(q.x, m2x[j], fmaf(q.y, m2y[j], q.w)) > lim[j]) continue;
        float lx, ly, lz;
        if (pp_local(b[j], q.x, q.y, q.z, lx, ly, lz)) {
          if (WRITE) {
            const int pos = before[j] + cnt[j];
            if (pos < max_inbox) scratch[r[j] * max_inbox + pos] = (int32_t)(t0 + i);
            else lim[j] = -INFINITY;
          }
          ++cnt[j];
        }
      }
    }
  }
  if (!WRITE) {
#pragma unroll
    for (int j = 0; j < kPpPerThread; ++j)
      if (r[j] < k) slice_counts[(int64_t)slice * k + r[j]] = cnt[j];
  }
}

// per roi: hits per slice -> exclusive prefix over the slices (in place), total -> counts (capped)
__global__ void __launch_bounds__(256)
    k_pp_slice_prefix(int32_t* __restrict__ slice_counts, int64_t k, int slices, int max_inbox, int32_t* __restrict__ counts) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= k) return;
  int run = 0;
  for (int s = 0; s < slices; ++s) {
    const int c = slice_counts[(int64_t)s * k + r];
    slice_counts[(int64_t)s * k + r] = run;
    run += c;
  }
  counts[r] = min(run, max_inbox);
}

// exclusive scan of the per-roi counts (k is a few thousand: one CTA), clipped to the output capacity
__global__ void __launch_bounds__(1024) k_pp_offsets(const int32_t* __restrict__ counts, int64_t k, int64_t capacity,
                                                     int32_t* __restrict__ offsets, int32_t* __restrict__ total) {
  __shared__ int s_warp[32];
  __shared__ int s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (int64_t base = 0; base < k; base += 1024) {
    const int64_t i = base + threadIdx.x;
    const int v = i < k ? counts[i] : 0;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, x, o);
      if ((int)lane_id() >= o) x += y;
    }
    if (lane_id() == 31) s_warp[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x < 32) {
      int t = s_warp[threadIdx.x];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, t, o);
        if ((int)threadIdx.x >= o) t += y;
      }
      s_warp[threadIdx.x] = t;
    }
    __syncthreads();
    const int before = s_carry + (threadIdx.x >= 32 ? s_warp[(threadIdx.x >> 5) - 1] : 0) + x - v;
    if (i < k) offsets[i] = (int32_t)min((int64_t)before, capacity);
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = before + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) total[0] = (int32_t)min((int64_t)s_carry, capacity);
}

__global__ void __launch_bounds__(kPpWarps * 32)
    k_pp_write(const float* __restrict__ rois, int64_t k, const float* __restrict__ pts, int64_t pts_stride, float e0,
               float e1, float e2, int max_inbox, const int32_t* __restrict__ scratch, const int32_t* __restrict__ counts,
               const int32_t* __restrict__ offsets, int64_t capacity, long long* __restrict__ out_pts_idx,
               long long* __restrict__ out_roi_idx, float* __restrict__ out_feats) {
  const int lane = lane_id();
  const int64_t r = (int64_t)blockIdx.x * kPpWarps + (threadIdx.x >> 5);
  if (r >= k) return;
  const PpBox b = pp_box(rois + r * 7, e0, e1, e2);
  const int cnt = counts[r];
  const int64_t off = offsets[r];
  for (int j = lane; j < cnt; j += 32) {
    const int64_t o = off + j;
    if (o >= capacity) break;
    const int32_t pi = scratch[r * max_inbox + j];
    const float* p = pts + (int64_t)pi * pts_stride;
    const float x = __ldg(p), y = __ldg(p + 1), z = __ldg(p + 2);
    float lx, ly, lz;
    pp_local(b, x, y, z, lx, ly, lz);
    const bool inner = (fabsf(lx) < b.hl) & (fabsf(ly) < b.hw) & (fabsf(lz) <= b.hh);
    out_pts_idx[o] = pi;
    out_roi_idx[o] = r;
    float* f = out_feats + o * 13;
    f[0] = x; f[1] = y; f[2] = z;
    f[3] = lx; f[4] = ly; f[5] = lz;
    f[6] = __fadd_rn(lx, b.hl); f[7] = __fadd_rn(ly, b.hw); f[8] = __fadd_rn(lz, b.hh);
    f[9] = __fsub_rn(b.hl, lx); f[10] = __fsub_rn(b.hw, ly); f[11] = __fsub_rn(b.hh, lz);
    f[12] = inner ? 0.f : 1.f;
  }
}

}  // namespace fsfb

namespace fsfb {
// BasePointBBoxCoder.decode (projects/mmdet3d_plugin/core/bbox/coders/base_point_bbox_coder.py:59-82) + the batch column of
// FSF.decode_stage_bboxes (models/detectors/FSF.py:1085-1095): reg = (dxyz, log dims, sin, cos[, vx, vy])
__global__ void __launch_bounds__(256)
    k_decode_boxes(const float* __restrict__ reg, int64_t k, int code, int64_t reg_stride, const float* __restrict__ base,
                   int64_t base_stride, const int32_t* __restrict__ batch, int64_t batch_stride, float eps, float* __restrict__ rois) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < k; i += (int64_t)gridDim.x * blockDim.x) {
    const float* r = reg + i * reg_stride;
    float* o = rois + i * (code);  // 1 + (code - 1) columns
    o[0] = batch ? (float)batch[i * batch_stride] : 0.f;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      o[1 + d] = __fadd_rn(__ldg(r + d), __ldg(base + i * base_stride + d));
      o[4 + d] = __fsub_rn(expf(__ldg(r + 3 + d)), eps);
    }
    o[7] = atan2f(__ldg(r + 6), __ldg(r + 7));
    if (code == 10) {
      o[8] = __ldg(r + 8);
      o[9] = __ldg(r + 9);
    }
  }
}
}  // namespace fsfb

extern "C" int fsfb_decode_boxes(const float* reg, int64_t k, int code_size, int64_t reg_stride, const float* base_points,
                                 int64_t base_stride, const int32_t* batch, int64_t batch_stride, float* rois, void* stream) {
  using namespace fsfb;
  FSFB_CHECK_ARG(k >= 0 && (code_size == 8 || code_size == 10) && reg_stride >= code_size && base_stride >= 3,
                 "decode_boxes: bad argument (code_size must be 8 or 10)");
  if (k == 0) return FSFB_OK;
  FSFB_CHECK_ARG(reg && base_points && rois, "decode_boxes: null pointer");
  const int grid = (int)std::min<int64_t>(ceil_div(k, 256), (int64_t)kNumSMs * 4);
  FSFB_LAUNCH(k_decode_boxes, grid, 256, 0, (cudaStream_t)stream, reg, k, code_size, reg_stride, base_points, base_stride, batch,
              batch_stride, 1e-6f, rois);
  return FSFB_OK;
}

extern "C" int fsfb_dynamic_point_pool_workspace_bytes(int64_t k, int max_inbox_point, size_t* bytes) {
  using namespace fsfb;
  FSFB_CHECK_ARG(bytes && k >= 0 && max_inbox_point >= 1, "dynamic_point_pool_workspace_bytes: bad argument");
  Workspace ws(nullptr, 0);
  ws.take<int32_t>((size_t)std::max<int64_t>(k, 1) * kPpMaxSlices);     // hits per (slice, roi)
  ws.take<int32_t>((size_t)std::max<int64_t>(k, 1) * max_inbox_point);  // scratch
  ws.take<int32_t>(std::max<int64_t>(k, 1));                             // counts
  ws.take<int32_t>(std::max<int64_t>(k, 1));                             // offsets
  *bytes = ws.used;
  return FSFB_OK;
}

extern "C" int fsfb_dynamic_point_pool(const float* rois, int64_t k, const float* pts, int64_t n, int64_t pts_stride,
                                       const float* extra_wlh, int max_inbox_point, int64_t capacity,
                                       long long* out_pts_idx, long long* out_roi_idx, float* out_pts_feats,
                                       int32_t* num_out, void* workspace, size_t workspace_bytes, void* stream) {
  using namespace fsfb;
  FSFB_CHECK_ARG(k >= 0 && n >= 0 && pts_stride >= 3 && max_inbox_point >= 1 && capacity >= 0 && k < (1ll << 31) &&
                     n < (1ll << 31),
                 "dynamic_point_pool: bad argument k=%lld n=%lld", (long long)k, (long long)n);
  FSFB_CHECK_ARG(extra_wlh && num_out, "dynamic_point_pool: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (k == 0 || n == 0 || capacity == 0) {
    FSFB_CUDA(cudaMemsetAsync(num_out, 0, sizeof(int32_t), st));
    return FSFB_OK;
  }
  FSFB_CHECK_ARG(rois && pts && out_pts_idx && out_roi_idx && out_pts_feats, "dynamic_point_pool: null device pointer");
  Workspace ws(workspace, workspace_bytes);
  int32_t* slice_counts = ws.take<int32_t>((size_t)k * kPpMaxSlices);
  int32_t* scratch = ws.take<int32_t>((size_t)k * max_inbox_point);
  int32_t* counts = ws.take<int32_t>(k);
  int32_t* offsets = ws.take<int32_t>(k);
  if (!ws.ok()) {
    set_error("dynamic_point_pool: workspace too small (%zu given, %zu needed)", workspace_bytes, ws.used);
    return FSFB_ERR_CAPACITY;
  }
  const int grid = (int)ceil_div(k, kPpWarps);
  // enough (roi block, slice) CTAs of two warps for ~24 warps per SM; slices are whole tiles
  const int rblocks = (int)ceil_div(k, kPpRois);
  int slices = (int)std::min<int64_t>(std::max<int64_t>(ceil_div(kNumSMs * 12, rblocks), 1), kPpMaxSlices);
  const int64_t slice_len = ceil_div(ceil_div(n, slices), kPpTile) * kPpTile;
  slices = (int)ceil_div(n, slice_len);
  const dim3 sgrid((unsigned)rblocks, (unsigned)slices);
  FSFB_LAUNCH(k_pp_scan<false>, sgrid, kPpThreads, 0, st, rois, k, pts, n, pts_stride, extra_wlh[0], extra_wlh[1], extra_wlh[2],
              max_inbox_point, slice_len, slice_counts, scratch);
  FSFB_LAUNCH(k_pp_slice_prefix, (int)ceil_div(k, 256), 256, 0, st, slice_counts, k, slices, max_inbox_point, counts);
  FSFB_LAUNCH(k_pp_scan<true>, sgrid, kPpThreads, 0, st, rois, k, pts, n, pts_stride, extra_wlh[0], extra_wlh[1], extra_wlh[2],
              max_inbox_point, slice_len, slice_counts, scratch);
  FSFB_LAUNCH(k_pp_offsets, 1, 1024, 0, st, counts, k, capacity, offsets, num_out);
  FSFB_LAUNCH(k_pp_write, grid, kPpWarps * 32, 0, st, rois, k, pts, pts_stride, extra_wlh[0], extra_wlh[1], extra_wlh[2],
              max_inbox_point, scratch, counts, offsets, capacity, out_pts_idx, out_roi_idx, out_pts_feats);
  return FSFB_OK;
}
