// pointwise.cu — the per-point elementwise / gather steps between the scatters and the GEMMs, each
// fused into one HBM pass (the reference spends ~10 ATen launches and several temporaries on each).
//   vfe_decorate   DynamicScatterVFE point decoration (un-vendored fork; config
//                  projects/configs/nuScenes/FSF_nuScenes_config.py:42-52): [f | xyz - mean | xyz - centre]
//   sir_input      SIRLayer input: cat(xyz / xyz_normalizer, feats) * rel_mlp gate (models/backbones/sir.py:41-62)
//   neck_points    Voxel2PointScatterNeck.forward (models/necks/voxel2point_neck.py:42-67)
//   vote_decode    VoteSegHead.decode_vote_targets (models/decode_heads/segmentation_head.py:265-266)
//   reduce_channel SparseUNet.reduce_channel (mmdet3d sparse_unet; used by SimpleSparseUNet's decoder)
//   compact_indices  boolean-mask compaction (extract_fg_pts FSF.py:299-308, group_sample single_stage_fsd.py:828-850)
// All HBM-bound; IEEE fp32 with no FMA contraction where the reference composes separate ATen ops.
#include "common.cuh"

namespace fsfb {

struct Vec3 {
  float x, y, z;
};

// out[i] = [ f[i, 0:cin] | xyz - mean[inv[i], 0:3] | xyz - (coor * vs + (vs/2 + min)) ]
template <typename CoorT>
__global__ void __launch_bounds__(256)
    k_vfe_decorate(const float* __restrict__ f, int64_t n, int cin, int64_t f_stride,
                   const CoorT* __restrict__ coors /* [n,4] b,z,y,x */, const int32_t* __restrict__ inv,
                   const float* __restrict__ mean /* [m, cin] */, Vec3 vs, Vec3 off, int with_cluster,
                   int with_center, float* __restrict__ out, int cout) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float* r = f + i * f_stride;
    float* o = out + i * cout;
    const float x = __ldg(r), y = __ldg(r + 1), z = __ldg(r + 2);
    for (int c = 0; c < cin; ++c) o[c] = __ldg(r + c);
    int p = cin;
    if (with_cluster) {
      const float* mu = mean + (int64_t)__ldg(inv + i) * cin;
      o[p] = __fsub_rn(x, __ldg(mu));
      o[p + 1] = __fsub_rn(y, __ldg(mu + 1));
      o[p + 2] = __fsub_rn(z, __ldg(mu + 2));
      p += 3;
    }
    if (with_center) {
      const CoorT* c4 = coors + i * 4;
      o[p] = __fsub_rn(x, __fadd_rn(__fmul_rn((float)c4[3], vs.x), off.x));
      o[p + 1] = __fsub_rn(y, __fadd_rn(__fmul_rn((float)c4[2], vs.y), off.y));
      o[p + 2] = __fsub_rn(z, __fadd_rn(__fmul_rn((float)c4[1], vs.z), off.z));
    }
  }
}

// out[i, c] = (c < 3 ? f[i,c] / nrm[c] : f[i,c]) * (gate ? gate[i,c] : 1)
__global__ void __launch_bounds__(256)
    k_sir_input(const float* __restrict__ f, int64_t n, int c, int64_t f_stride, Vec3 nrm,
                const float* __restrict__ gate, int64_t gate_stride, float* __restrict__ out, int64_t out_stride) {
  const int64_t total = n * c;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = t / c;
    const int j = (int)(t - i * c);
    float v = __ldg(f + i * f_stride + j);
    if (j == 0) v = __fdiv_rn(v, nrm.x);
    if (j == 1) v = __fdiv_rn(v, nrm.y);
    if (j == 2) v = __fdiv_rn(v, nrm.z);
    if (gate) v = __fmul_rn(v, __ldg(gate + i * gate_stride + j));
    out[i * out_stride + j] = v;
  }
}

// out[i, c] = x[i, c] / d[c]     (f_cluster / rel_dist_scaler)
__global__ void __launch_bounds__(256)
    k_div_cols(const float* __restrict__ x, int64_t n, int c, int64_t x_stride, const float* __restrict__ d,
               float* __restrict__ out, int64_t out_stride) {
  const int64_t total = n * c;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = t / c;
    const int j = (int)(t - i * c);
    out[i * out_stride + j] = __fdiv_rn(__ldg(x + i * x_stride + j), __ldg(d + j));
  }
}

__global__ void __launch_bounds__(256)
    k_add_inplace(float* __restrict__ x, int64_t n, int c, int64_t x_stride, const float* __restrict__ y,
                  int64_t y_stride) {
  const int64_t total = n * c;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = t / c;
    const int j = (int)(t - i * c);
    x[i * x_stride + j] = __fadd_rn(x[i * x_stride + j], __ldg(y + i * y_stride + j));
  }
}

// out[i, c] = sum_{j < r} x[i, c*r + j]   (features.view(n, out_channels, -1).sum(dim=2))
__global__ void __launch_bounds__(256)
    k_reduce_channel(const float* __restrict__ x, int64_t n, int cin, int64_t x_stride, int cout,
                     float* __restrict__ out) {
  const int r = cin / cout;
  const int64_t total = n * cout;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = t / cout;
    const int c = (int)(t - i * cout);
    const float* p = x + i * x_stride + (int64_t)c * r;
    float s = __ldg(p);
    for (int j = 1; j < r; ++j) s = __fadd_rn(s, __ldg(p + j));
    out[t] = s;
  }
}

// Voxel2PointScatterNeck: out[i] = [ vf[inv[i], :] | xyz - ((coor + 0.5) * vs + min) ], mask[i] = row != padding.
// One warp per point row (lanes over channels, 128-bit when C % 4 == 0).
template <typename CoorT, typename IdxT>
__global__ void __launch_bounds__(256)
    k_neck_points(const float* __restrict__ pts, int64_t n, int64_t pts_stride, const CoorT* __restrict__ coors,
                  const float* __restrict__ vf, int64_t m, int c, const IdxT* __restrict__ inv, Vec3 vs, Vec3 lo,
                  float padding, float* __restrict__ out, int64_t out_stride, uint8_t* __restrict__ mask,
                  int* __restrict__ dropped) {
  const int lane = lane_id();
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < n; i += warps) {
    const int64_t s = (int64_t)inv[i];
    const float* src = vf + s * c;
    float* o = out + i * out_stride;
    bool all_pad = true;
    for (int j = lane; j < c; j += 32) {
      const float v = __ldg(src + j);
      all_pad &= (v == padding);
      o[j] = v;
    }
    all_pad = __all_sync(0xffffffffu, all_pad);
    if (lane < 3) {
      const float p = __ldg(pts + i * pts_stride + lane);
      const float cc = (float)coors[i * 4 + (3 - lane)];  // x←col 3, y←col 2, z←col 1
      const float v = lane == 0 ? vs.x : (lane == 1 ? vs.y : vs.z);
      const float l = lane == 0 ? lo.x : (lane == 1 ? lo.y : lo.z);
      o[c + lane] = __fsub_rn(p, __fadd_rn(__fmul_rn(__fadd_rn(cc, 0.5f), v), l));
    }
    if (lane == 0) {
      mask[i] = all_pad ? 0 : 1;
      if (all_pad) atomicAdd(dropped, 1);
    }
  }
}

__global__ void __launch_bounds__(256) k_vote_decode(const float* __restrict__ x, int64_t total, float* __restrict__ out) {
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const float v = __ldg(x + t);
    out[t] = __fmul_rn(v, fabsf(v));
  }
}

// ---- stable mask compaction: idx[k] = position of the k-th set flag ---------------------------
constexpr int kCpThreads = 256;
constexpr int kCpItems = 8;
constexpr int kCpTile = kCpThreads * kCpItems;

__device__ __forceinline__ int cta_excl_scan_i32(int v, int* s_warp /*[32]*/, int* total) {
  int x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int y = __shfl_up_sync(0xffffffffu, x, o);
    if ((int)lane_id() >= o) x += y;
  }
  const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (lane_id() == 31) s_warp[w] = x;
  __syncthreads();
  if (w == 0) {
    const int s = (int)lane_id() < nw ? s_warp[lane_id()] : 0;
    int t = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, t, o);
      if ((int)lane_id() >= o) t += y;
    }
    s_warp[lane_id()] = t - s;
    if (lane_id() == 31 && total) *total = t;
  }
  __syncthreads();
  const int r = s_warp[w] + x - v;
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(kCpThreads)
    k_compact_count(const uint8_t* __restrict__ mask, int64_t n, int* __restrict__ tile_sums) {
  __shared__ int sw[32];
  __shared__ int total;
  const int64_t base = (int64_t)blockIdx.x * kCpTile + (int64_t)threadIdx.x * kCpItems;
  int s = 0;
#pragma unroll
  for (int k = 0; k < kCpItems; ++k)
    if (base + k < n) s += mask[base + k] != 0;
  cta_excl_scan_i32(s, sw, &total);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) k_compact_scan(int* __restrict__ tile_sums, int ntiles, int* __restrict__ count) {
  __shared__ int sw[32];
  __shared__ int total;
  int carry = 0;
  for (int base = 0; base < ntiles; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = i < ntiles ? tile_sums[i] : 0;
    const int ex = cta_excl_scan_i32(v, sw, &total);
    if (i < ntiles) tile_sums[i] = carry + ex;
    carry += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) *count = carry;
}

__global__ void __launch_bounds__(kCpThreads)
    k_compact_write(const uint8_t* __restrict__ mask, int64_t n, const int* __restrict__ tile_sums,
                    int32_t* __restrict__ idx) {
  __shared__ int sw[32];
  const int64_t base = (int64_t)blockIdx.x * kCpTile + (int64_t)threadIdx.x * kCpItems;
  int f[kCpItems], s = 0;
#pragma unroll
  for (int k = 0; k < kCpItems; ++k) {
    f[k] = (base + k < n) ? (mask[base + k] != 0) : 0;
    s += f[k];
  }
  int ex = cta_excl_scan_i32(s, sw, nullptr) + tile_sums[blockIdx.x];
#pragma unroll
  for (int k = 0; k < kCpItems; ++k)
    if (f[k]) idx[ex++] = (int32_t)(base + k);
}

// mask[i] = x[i*stride + col] > thr     (fg_mask = seg_scores > cls_score_thr, single_stage_fsd.py:756)
__global__ void __launch_bounds__(256)
    k_threshold_mask(const float* __restrict__ x, int64_t n, int64_t stride, int col, float thr, uint8_t* __restrict__ mask) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    mask[i] = __ldg(x + i * stride + col) > thr ? 1 : 0;
}

// mask[i] = counts[inv[i]] >= min_count   (filter_almost_empty, single_stage_fsd.py:31-35)
__global__ void __launch_bounds__(256)
    k_count_mask(const int32_t* __restrict__ counts, const int32_t* __restrict__ inv, int64_t n, int min_count,
                 uint8_t* __restrict__ mask) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    mask[i] = __ldg(counts + __ldg(inv + i)) >= min_count ? 1 : 0;
}

static int grid1d(int64_t total, int per_sm = 16) {
  return (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(total, 256), (int64_t)kNumSMs * per_sm));
}

}  // namespace fsfb

extern "C" {

int fsfb_vfe_decorate(const float* feats, int64_t n, int cin, int64_t feat_stride, const void* coors,
                      int coors_i64, const int32_t* inv, const float* voxel_mean, const float* voxel_size,
                      const float* range_min, int with_cluster_center, int with_voxel_center, float* out,
                      void* stream) {
  using namespace fsfb;
  FSFB_CHECK_ARG(n >= 0 && cin >= 3 && feat_stride >= cin && voxel_size && range_min, "vfe_decorate: bad argument");
  if (n == 0) return FSFB_OK;
  FSFB_CHECK_ARG(feats && out && (!with_voxel_center || coors) && (!with_cluster_center || (inv && voxel_mean)),
                 "vfe_decorate: null pointer");
  const int cout = cin + 3 * (with_cluster_center != 0) + 3 * (with_voxel_center != 0);
  Vec3 vs{voxel_size[0], voxel_size[1], voxel_size[2]};
  // x_offset = vx / 2 + range_min evaluated in double (Python floats), then rounded to fp32 when it meets the tensor
  Vec3 off{(float)((double)voxel_size[0] / 2 + (double)range_min[0]), (float)((double)voxel_size[1] / 2 + (double)range_min[1]),
           (float)((double)voxel_size[2] / 2 + (double)range_min[2])};
  cudaStream_t st = (cudaStream_t)stream;
  if (coors_i64) {
    FSFB_LAUNCH(k_vfe_decorate<long long>, grid1d(n), 256, 0, st, feats, n, cin, feat_stride, (const long long*)coors,
                inv, voxel_mean, vs, off, with_cluster_center, with_voxel_center, out, cout);
  } else {
    FSFB_LAUNCH(k_vfe_decorate<int>, grid1d(n), 256, 0, st, feats, n, cin, feat_stride, (const int*)coors, inv,
                voxel_mean, vs, off, with_cluster_center, with_voxel_center, out, cout);
  }
  return FSFB_OK;
}

int fsfb_sir_input(const float* feats, int64_t n, int c, int64_t feat_stride, const float* xyz_normalizer,
                   const float* gate, int64_t gate_stride, float* out, int64_t out_stride, void* stream) {
  using namespace fsfb;
  FSFB_CHECK_ARG(n >= 0 && c >= 3 && feat_stride >= c && out_stride >= c && xyz_normalizer, "sir_input: bad argument");
  if (n == 0) return FSFB_OK;
  FSFB_CHECK_ARG(feats && out, "sir_input: null pointer");
  Vec3 nrm{xyz_normalizer[0], xyz_normalizer[1], xyz_normalizer[2]};
  FSFB_LAUNCH(k_sir_input, grid1d(n * c), 256, 0, (cudaStream_t)stream, feats, n, c, feat_stride, nrm, gate, gate_stride,
              out, out_stride);
  return FSFB_OK;
}

int fsfb_div_cols(const float* x, int64_t n, int c, int64_t x_stride, const float* divisors_dev, float* out,
                  int64_t out_stride, void* stream) {
  using namespace fsfb;
  FSFB_CHECK_ARG(n >= 0 && c >= 1 && x_stride >= c && out_stride >= c, "div_cols: bad argument");
  if (n == 0) return FSFB_OK;
  FSFB_CHECK_ARG(x && out && divisors_dev, "div_cols: null pointer");
  FSFB_LAUNCH(k_div_cols, grid1d(n * c), 256, 0, (cudaStream_t)stream, x, n, c, x_stride, divisors_dev, out, out_stride);
  return FSFB_OK;
}

int fsfb_add_inplace(float* x, int64_t n, int c, int64_t x_stride, const float* y, int64_t y_stride, void* stream) {
  using namespace fsfb;
  FSFB_CHECK_ARG(n >= 0 && c >= 1 && x_stride >= c && y_stride >= c, "add_inplace: bad argument");
  if (n == 0) return FSFB_OK;
  FSFB_CHECK_ARG(x && y, "add_inplace: null pointer");
  FSFB_LAUNCH(k_add_inplace, grid1d(n * c), 256, 0, (cudaStream_t)stream, x, n, c, x_stride, y, y_stride);
  return FSFB_OK;
}

int fsfb_reduce_channel(const float* x, int64_t n, int cin, int64_t x_stride, int cout, float* out, void* stream) {
  using namespace fsfb;
  FSFB_CHECK_ARG(n >= 0 && cout >= 1 && cin >= cout && cin % cout == 0 && x_stride >= cin, "reduce_channel: bad shape");
  if (n == 0) return FSFB_OK;
  FSFB_CHECK_ARG(x && out, "reduce_channel: null pointer");
  FSFB_LAUNCH(k_reduce_channel, grid1d(n * cout), 256, 0, (cudaStream_t)stream, x, n, cin, x_stride, cout, out);
  return FSFB_OK;
}

int fsfb_neck_points(const float* points, int64_t n, int64_t pts_stride, const void* coors, int coors_i64,
                     const float* voxel_feats, int64_t m, int c, const void* inv, int inv_i64,
                     const float* voxel_size, const float* range_min, float padding, float* out, int64_t out_stride,
                     uint8_t* mask, int32_t* dropped, void* stream) {
  using namespace fsfb;
  FSFB_CHECK_ARG(n >= 0 && m >= 0 && c >= 1 && pts_stride >= 3 && out_stride >= c + 3 && voxel_size && range_min,
                 "neck_points: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (dropped) FSFB_CUDA(cudaMemsetAsync(dropped, 0, 4, st));
  if (n == 0) return FSFB_OK;
  FSFB_CHECK_ARG(points && coors && voxel_feats && inv && out && mask && dropped, "neck_points: null pointer");
  Vec3 vs{voxel_size[0], voxel_size[1], voxel_size[2]}, lo{range_min[0], range_min[1], range_min[2]};
  const int grid = (int)std::min<int64_t>(ceil_div(n, 8), (int64_t)kNumSMs * 32);
#define NECK(CT, IT)                                                                                          \
  FSFB_LAUNCH((k_neck_points<CT, IT>), grid, 256, 0, st, points, n, pts_stride, (const CT*)coors, voxel_feats, m, c, \
              (const IT*)inv, vs, lo, padding, out, out_stride, mask, dropped)
  if (coors_i64 && inv_i64) NECK(long long, long long);
  else if (coors_i64) NECK(long long, int);
  else if (inv_i64) NECK(int, long long);
  else NECK(int, int);
#undef NECK
  return FSFB_OK;
}

int fsfb_vote_decode(const float* preds, int64_t total, float* out, void* stream) {
  using namespace fsfb;
  FSFB_CHECK_ARG(total >= 0, "vote_decode: bad size");
  if (total == 0) return FSFB_OK;
  FSFB_CHECK_ARG(preds && out, "vote_decode: null pointer");
  FSFB_LAUNCH(k_vote_decode, grid1d(total), 256, 0, (cudaStream_t)stream, preds, total, out);
  return FSFB_OK;
}

int fsfb_threshold_mask(const float* x, int64_t n, int64_t stride, int col, float thr, uint8_t* mask, void* stream) {
  using namespace fsfb;
  FSFB_CHECK_ARG(n >= 0 && col >= 0 && col < stride, "threshold_mask: bad argument");
  if (n == 0) return FSFB_OK;
  FSFB_CHECK_ARG(x && mask, "threshold_mask: null pointer");
  FSFB_LAUNCH(k_threshold_mask, grid1d(n), 256, 0, (cudaStream_t)stream, x, n, stride, col, thr, mask);
  return FSFB_OK;
}

int fsfb_count_mask(const int32_t* counts, const int32_t* inv, int64_t n, int min_count, uint8_t* mask, void* stream) {
  using namespace fsfb;
  FSFB_CHECK_ARG(n >= 0, "count_mask: bad argument");
  if (n == 0) return FSFB_OK;
  FSFB_CHECK_ARG(counts && inv && mask, "count_mask: null pointer");
  FSFB_LAUNCH(k_count_mask, grid1d(n), 256, 0, (cudaStream_t)stream, counts, inv, n, min_count, mask);
  return FSFB_OK;
}

int fsfb_compact_workspace_bytes(int64_t n, size_t* bytes) {
  using namespace fsfb;
  FSFB_CHECK_ARG(bytes && n >= 0 && n < (1ll << 31), "compact_workspace_bytes: bad argument");
  *bytes = align_up((size_t)std::max<int64_t>(1, ceil_div(n, kCpTile)) * 4, 256);
  return FSFB_OK;
}

int fsfb_compact_indices(const uint8_t* mask, int64_t n, int32_t* idx, int32_t* count, void* workspace,
                         size_t workspace_bytes, void* stream) {
  using namespace fsfb;
  FSFB_CHECK_ARG(n >= 0 && n < (1ll << 31) && count, "compact_indices: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0) {
    FSFB_CUDA(cudaMemsetAsync(count, 0, 4, st));
    return FSFB_OK;
  }
  FSFB_CHECK_ARG(mask && idx, "compact_indices: null pointer");
  const int ntiles = (int)ceil_div(n, kCpTile);
  if (!workspace || workspace_bytes < (size_t)ntiles * 4) {
    set_error("compact_indices: workspace too small");
    return FSFB_ERR_CAPACITY;
  }
  int* tile_sums = (int*)workspace;
  FSFB_LAUNCH(k_compact_count, ntiles, kCpThreads, 0, st, mask, n, tile_sums);
  FSFB_LAUNCH(k_compact_scan, 1, 1024, 0, st, tile_sums, ntiles, (int*)count);
  FSFB_LAUNCH(k_compact_write, ntiles, kCpThreads, 0, st, mask, n, tile_sums, idx);
  return FSFB_OK;
}

}  // extern "C"
