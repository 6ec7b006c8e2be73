// project.cu — LiDAR→camera projection fused with nearest sampling of the 2D instance-id planes
// (a7 + a8 + the camera selection of a9).
//
// Reference semantics, op for op: FSF.prj_points_2d + FSF.points_in_mask
// (projects/mmdet3d_plugin/models/detectors/FSF.py:169-226) and the camera selection of
// FSF.img_cross_attn (:714-718) / FSF.extract_fg_pts (:299-308):
//   pts_2d = [x,y,z,1] @ lidar2img^T;  depth_valid = z' > 1e-3;  z' = clip(z', 1e-5, 1e5)
//   u = x'/z' / W;  v = y'/z' / H;  g = (uv - 0.5) * 2;  valid = depth_valid & g in (-1,1)^2
//   invalid → g = -2;  id = grid_sample(mask.float(), g, mode='nearest', align_corners=False)
// grid_sample(nearest): ix = ((g+1)*W - 1)/2 (the CUDA kernel's form, with the multiply-subtract
// contracted to one FMA as nvcc compiles ATen), texel = nearbyint(ix) (half to even), zeros
// outside.  The reference first casts the whole [cams,classes,H,W] u8 tensor to f32 (346 MB
// written per call at nuScenes size); here the u8 / i32 planes are sampled as stored.
//
// HBM-bound.  Algorithmic bytes per point: 12 (xyz) + cams*classes*1 (texels) +
// cams*classes*8 (i64 ids) = 552 B for the drop-in contract at 6x10; 12 + 60 + 4*classes + 3
// for the fused contract.  Drop-in kernel: ids are staged in shared memory per warp and the
// [32 pts x cams x classes] i64 block is written as one contiguous run of 16-byte stores.
#include "common.cuh"

namespace fsfb {

constexpr int kMaxCams = 8;
constexpr int kMaxClasses = 16;

struct CamSet {
  float P[kMaxCams][12];  // first three rows of each 4x4 lidar2img
};

__device__ __forceinline__ void load_cams(CamSet& s, const float* __restrict__ lidar2img, int cams) {
  for (int t = threadIdx.x; t < cams * 12; t += blockDim.x)
    s.P[t / 12][t % 12] = __ldg(lidar2img + (t / 12) * 16 + (t % 12));
  __syncthreads();
}

// Returns texel offset (iy*W+ix) or -1 when the point does not sample this camera.
__device__ __forceinline__ int project_texel(const float* __restrict__ P, float x, float y, float z,
                                             int W, int H) {
  // [x y z 1] . row_j — sequential FMA accumulation over k (K = 4 GEMM)
  float xc = __fadd_rn(fmaf(z, P[2], fmaf(y, P[1], __fmul_rn(x, P[0]))), P[3]);
  float yc = __fadd_rn(fmaf(z, P[6], fmaf(y, P[5], __fmul_rn(x, P[4]))), P[7]);
  float zc = __fadd_rn(fmaf(z, P[10], fmaf(y, P[9], __fmul_rn(x, P[8]))), P[11]);
  const bool depth_ok = zc > 1e-3f;
  zc = fminf(fmaxf(zc, 1e-5f), 1e5f);
  float u = __fdiv_rn(__fdiv_rn(xc, zc), (float)W);
  float v = __fdiv_rn(__fdiv_rn(yc, zc), (float)H);
  float gx = __fmul_rn(__fsub_rn(u, 0.5f), 2.f);
  float gy = __fmul_rn(__fsub_rn(v, 0.5f), 2.f);
  const bool ok = depth_ok & (gx > -1.f) & (gx < 1.f) & (gy > -1.f) & (gy < 1.f);
  if (!ok) return -1;
  // grid_sampler_unnormalize (align_corners = False) + nearbyint
  float ix = __fdiv_rn(fmaf(__fadd_rn(gx, 1.f), (float)W, -1.f), 2.f);
  float iy = __fdiv_rn(fmaf(__fadd_rn(gy, 1.f), (float)H, -1.f), 2.f);
  int ixn = __float2int_rn(ix), iyn = __float2int_rn(iy);
  if (ixn < 0 || ixn >= W || iyn < 0 || iyn >= H) return -1;
  return iyn * W + ixn;
}

// ---- drop-in contract: out_ids [n, cams, classes] i64 ----------------------------------
template <typename MaskT>
__global__ void __launch_bounds__(256)
    k_project_sample(const float* __restrict__ xyz, int64_t n, int64_t stride,
                     const float* __restrict__ lidar2img, int cams, const MaskT* __restrict__ mask,
                     int classes, int H, int W, long long* __restrict__ out) {
  extern __shared__ unsigned char s_raw[];
  __shared__ CamSet s_cams;
  load_cams(s_cams, lidar2img, cams);
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  const int per_pt = cams * classes;
  MaskT* s_ids = reinterpret_cast<MaskT*>(s_raw) + (size_t)warp * 32 * per_pt;
  const int64_t plane = (int64_t)H * W;
  const int64_t n_warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t base = ((int64_t)blockIdx.x * (blockDim.x >> 5) + warp) * 32; base < n;
       base += n_warps * 32) {
    const int64_t i = base + lane;
    if (i < n) {
      const float* p = xyz + i * stride;
      const float x = __ldg(p), y = __ldg(p + 1), z = __ldg(p + 2);
      MaskT* mine = s_ids + lane * per_pt;
      for (int cam = 0; cam < cams; ++cam) {
        const int tex = project_texel(s_cams.P[cam], x, y, z, W, H);
        if (tex >= 0) {
          const MaskT* m0 = mask + (int64_t)cam * classes * plane + tex;
#pragma unroll 10
          for (int k = 0; k < classes; ++k) mine[cam * classes + k] = __ldg(m0 + k * plane);
        } else {
          for (int k = 0; k < classes; ++k) mine[cam * classes + k] = 0;
        }
      }
    }
    __syncwarp();
    // coalesced i64 write of the warp's contiguous block
    const int64_t pts_here = min((int64_t)32, n - base);
    const int64_t total = pts_here * per_pt;  // i64 elements
    long long* o = out + base * per_pt;
    if ((total & 1) == 0 && ((uintptr_t)o & 15) == 0) {
      for (int64_t e = (int64_t)lane * 2; e < total; e += 64) {
        longlong2 v = make_longlong2((long long)s_ids[e], (long long)s_ids[e + 1]);
        *reinterpret_cast<longlong2*>(o + e) = v;
      }
    } else {
      for (int64_t e = lane; e < total; e += 32) o[e] = (long long)s_ids[e];
    }
    __syncwarp();
  }
}

// ---- fused contract: camera-selected ids + fg flag --------------------------------------
template <typename MaskT>
__global__ void __launch_bounds__(256)
    k_project_sample_select(const float* __restrict__ xyz, int64_t n, int64_t stride,
                            const float* __restrict__ lidar2img, int cams,
                            const MaskT* __restrict__ mask, int classes, int H, int W,
                            int32_t* __restrict__ ids_sel, uint8_t* __restrict__ cam_sel,
                            uint8_t* __restrict__ fg, uint8_t* __restrict__ overlap,
                            const float* __restrict__ anno, int anno_rows, int anno_cols, int anno_col,
                            float* __restrict__ scores) {
  __shared__ CamSet s_cams;
  load_cams(s_cams, lidar2img, cams);
  const int64_t plane = (int64_t)H * W;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const float* p = xyz + i * stride;
    const float x = __ldg(p), y = __ldg(p + 1), z = __ldg(p + 2);
    int best_ids[kMaxClasses];
#pragma unroll
    for (int k = 0; k < kMaxClasses; ++k) best_ids[k] = 0;
    long long best_sum = 0;  // an all-zero camera 0 wins ties, as torch.max returns the first maximum
    int best_cam = 0, n_pos = 0;
    for (int cam = 0; cam < cams; ++cam) {
      const int tex = project_texel(s_cams.P[cam], x, y, z, W, H);
      if (tex < 0) continue;
      const MaskT* m0 = mask + (int64_t)cam * classes * plane + tex;
      int ids[kMaxClasses];
      long long sum = 0;
#pragma unroll
      for (int k = 0; k < kMaxClasses; ++k) {
        ids[k] = (k < classes) ? (int)__ldg(m0 + k * plane) : 0;
        sum += ids[k];
        n_pos += ids[k] > 0;
      }
      if (sum > best_sum) {
        best_sum = sum;
        best_cam = cam;
#pragma unroll
        for (int k = 0; k < kMaxClasses; ++k) best_ids[k] = ids[k];
      }
    }
    if (ids_sel) {
      int32_t* o = ids_sel + i * classes;
#pragma unroll
      for (int k = 0; k < kMaxClasses; ++k)
        if (k < classes) o[k] = best_ids[k];
    }
    if (cam_sel) cam_sel[i] = (uint8_t)best_cam;
    if (fg) fg[i] = (uint8_t)(best_sum > 0 || n_pos > 0);
    if (overlap) overlap[i] = (uint8_t)min(n_pos, 255);
    if (scores) {
      // get_all_cls_preds_2d + encode_preds_2d(encode_single_cls=False) (FSF.py:506-535, 449-474):
      // column `anno_col` (the 2D score) of mask_anno[id - 1]; id == 0 → 0
      float* so = scores + i * classes;
#pragma unroll
      for (int k = 0; k < kMaxClasses; ++k)
        if (k < classes) {
          const int id = best_ids[k];
          so[k] = (id >= 1 && id <= anno_rows) ? __ldg(anno + (int64_t)(id - 1) * anno_cols + anno_col) : 0.f;
        }
    }
  }
}

// ---- fused contract on class-interleaved planes (EXPERIMENTAL, fsfb_project_sample_select_hwc) -----------------------------
// mask [cams, H, W, 16] u8: the ids of a texel are ONE aligned 16-byte load instead of `classes` loads from as many planes
// (= sectors).  The [32 points x classes] id and score blocks of a warp are staged in shared memory and leave as contiguous
// 128-byte stores (the per-thread form writes 4 bytes every 40: ten partial-sector writes per point and output).
__global__ void __launch_bounds__(256)
    k_project_sample_select_hwc(const float* __restrict__ xyz, int64_t n, int64_t stride, const float* __restrict__ lidar2img, int cams,
                                const uint4* __restrict__ mask, int classes, int H, int W, int32_t* __restrict__ ids_sel,
                                uint8_t* __restrict__ cam_sel, uint8_t* __restrict__ fg, uint8_t* __restrict__ overlap,
                                const float* __restrict__ anno, int anno_rows, int anno_cols, int anno_col,
                                float* __restrict__ scores) {
  static_assert(kMaxClasses == 16, "one 16-byte texel");
  __shared__ CamSet s_cams;
  __shared__ int32_t s_ids[8][32 * kMaxClasses];
  __shared__ float s_sc[8][32 * kMaxClasses];
  load_cams(s_cams, lidar2img, cams);
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  const int64_t plane = (int64_t)H * W;
  for (int64_t base = (int64_t)blockIdx.x * blockDim.x + warp * 32; base < n; base += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = base + lane;
    if (i < n) {
      const float* p = xyz + i * stride;
      const float x = __ldg(p), y = __ldg(p + 1), z = __ldg(p + 2);
      uint4 best = make_uint4(0u, 0u, 0u, 0u);
      int best_sum = 0, best_cam = 0, n_pos = 0;  // an all-zero camera 0 wins ties, as torch.max returns the first maximum
      for (int cam = 0; cam < cams; ++cam) {
        const int tex = project_texel(s_cams.P[cam], x, y, z, W, H);
        if (tex < 0) continue;
        const uint4 t = __ldg(mask + ((int64_t)cam * plane + tex));
        const uint32_t w4[4] = {t.x, t.y, t.z, t.w};
        int sum = 0;
#pragma unroll
        for (int k = 0; k < kMaxClasses; ++k) {
          const int id = (k < classes) ? (int)((w4[k >> 2] >> (8 * (k & 3))) & 0xffu) : 0;
          sum += id;
          n_pos += id > 0;
        }
        if (sum > best_sum) {
          best_sum = sum;
          best_cam = cam;
          best = t;
        }
      }
      const uint32_t b4[4] = {best.x, best.y, best.z, best.w};
#pragma unroll
      for (int k = 0; k < kMaxClasses; ++k)
        if (k < classes) {
          const int id = (int)((b4[k >> 2] >> (8 * (k & 3))) & 0xffu);
          s_ids[warp][lane * classes + k] = id;
          if (scores) s_sc[warp][lane * classes + k] = (id >= 1 && id <= anno_rows) ? __ldg(anno + (int64_t)(id - 1) * anno_cols + anno_col) : 0.f;
        }
      if (cam_sel) cam_sel[i] = (uint8_t)best_cam;
      if (fg) fg[i] = (uint8_t)(best_sum > 0 || n_pos > 0);
      if (overlap) overlap[i] = (uint8_t)min(n_pos, 255);
    }
    __syncwarp();
    const int total = (int)min((int64_t)32, n - base) * classes;
    if (ids_sel)
      for (int e = lane; e < total; e += 32) ids_sel[base * classes + e] = s_ids[warp][e];
    if (scores)
      for (int e = lane; e < total; e += 32) scores[base * classes + e] = s_sc[warp][e];
    __syncwarp();
  }
}

static int check_common(const float* xyz, int64_t n, int64_t xyz_stride, const float* lidar2img,
                        int cams, const void* mask, int classes, int H, int W, const char* who) {
  FSFB_CHECK_ARG(n >= 0 && xyz_stride >= 3, "%s: bad n/stride", who);
  FSFB_CHECK_ARG(cams >= 1 && cams <= kMaxCams, "%s: cams=%d unsupported (1..%d)", who, cams, kMaxCams);
  FSFB_CHECK_ARG(classes >= 1 && classes <= kMaxClasses, "%s: classes=%d unsupported (1..%d)", who,
                 classes, kMaxClasses);
  FSFB_CHECK_ARG(H >= 1 && W >= 1 && (int64_t)H * W < (1ll << 31), "%s: bad H/W", who);
  FSFB_CHECK_ARG(n == 0 || (xyz && lidar2img && mask), "%s: null pointer", who);
  return FSFB_OK;
}

}  // namespace fsfb

extern "C" {

int fsfb_project_sample(const float* xyz, int64_t n, int64_t xyz_stride, const float* lidar2img,
                        int cams, const void* mask, int mask_i32, int classes, int H, int W,
                        int64_t* out_ids, void* stream) {
  using namespace fsfb;
  int rc = check_common(xyz, n, xyz_stride, lidar2img, cams, mask, classes, H, W, "project_sample");
  if (rc != FSFB_OK) return rc;
  if (n == 0) return FSFB_OK;
  FSFB_CHECK_ARG(out_ids, "project_sample: null output");
  cudaStream_t st = (cudaStream_t)stream;
  const int threads = 256;
  const size_t smem = (size_t)(threads / 32) * 32 * cams * classes * (mask_i32 ? 4 : 1);
  const int grid = (int)std::min<int64_t>(ceil_div(n, threads), (int64_t)kNumSMs * 8);
  if (mask_i32) {
    static bool attr = false;
    if (!attr) {
      FSFB_CUDA(cudaFuncSetAttribute(k_project_sample<int>, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072));
      attr = true;
    }
    FSFB_LAUNCH(k_project_sample<int>, grid, threads, smem, st, xyz, n, xyz_stride, lidar2img, cams,
                (const int*)mask, classes, H, W, (long long*)out_ids);
  } else {
    FSFB_LAUNCH(k_project_sample<unsigned char>, grid, threads, smem, st, xyz, n, xyz_stride, lidar2img, cams,
                (const unsigned char*)mask, classes, H, W, (long long*)out_ids);
  }
  return FSFB_OK;
}

int fsfb_project_sample_select(const float* xyz, int64_t n, int64_t xyz_stride,
                               const float* lidar2img, int cams, const void* mask, int mask_i32,
                               int classes, int H, int W, int32_t* ids_sel, uint8_t* cam_sel,
                               uint8_t* fg, uint8_t* overlap, const float* anno, int anno_rows,
                               int anno_cols, int anno_col, float* scores, void* stream) {
  using namespace fsfb;
  int rc = check_common(xyz, n, xyz_stride, lidar2img, cams, mask, classes, H, W,
                        "project_sample_select");
  if (rc != FSFB_OK) return rc;
  if (n == 0) return FSFB_OK;
  FSFB_CHECK_ARG(ids_sel || cam_sel || fg || overlap || scores, "project_sample_select: no output requested");
  FSFB_CHECK_ARG(!scores || (anno && anno_rows >= 0 && anno_col >= 0 && anno_col < anno_cols),
                 "project_sample_select: scores need a valid annotation table");
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = (int)std::min<int64_t>(ceil_div(n, 256), (int64_t)kNumSMs * 8);
  if (mask_i32) {
    FSFB_LAUNCH(k_project_sample_select<int>, grid, 256, 0, st, xyz, n, xyz_stride, lidar2img, cams,
                (const int*)mask, classes, H, W, ids_sel, cam_sel, fg, overlap, anno, anno_rows, anno_cols,
                anno_col, scores);
  } else {
    FSFB_LAUNCH(k_project_sample_select<unsigned char>, grid, 256, 0, st, xyz, n, xyz_stride, lidar2img, cams,
                (const unsigned char*)mask, classes, H, W, ids_sel, cam_sel, fg, overlap, anno, anno_rows,
                anno_cols, anno_col, scores);
  }
  return FSFB_OK;
}

// Experimental twin of fsfb_project_sample_select for class-interleaved u8 planes: mask dev [cams, H, W, 16] u8 (texel = 16 bytes,
// byte k = id of class k, bytes >= classes ignored), 16-byte aligned.  Same outputs.
int fsfb_project_sample_select_hwc(const float* xyz, int64_t n, int64_t xyz_stride, const float* lidar2img, int cams,
                                   const void* mask_hwc16, int classes, int H, int W, int32_t* ids_sel, uint8_t* cam_sel,
                                   uint8_t* fg, uint8_t* overlap, const float* anno, int anno_rows, int anno_cols, int anno_col,
                                   float* scores, void* stream) {
  using namespace fsfb;
  int rc = check_common(xyz, n, xyz_stride, lidar2img, cams, mask_hwc16, classes, H, W, "project_sample_select_hwc");
  if (rc != FSFB_OK) return rc;
  if (n == 0) return FSFB_OK;
  FSFB_CHECK_ARG(((uintptr_t)mask_hwc16 & 15) == 0, "project_sample_select_hwc: planes must be 16-byte aligned");
  FSFB_CHECK_ARG(ids_sel || cam_sel || fg || overlap || scores, "project_sample_select_hwc: no output requested");
  FSFB_CHECK_ARG(!scores || (anno && anno_rows >= 0 && anno_col >= 0 && anno_col < anno_cols),
                 "project_sample_select_hwc: scores need a valid annotation table");
  const int grid = (int)std::min<int64_t>(ceil_div(n, 256), (int64_t)kNumSMs * 8);
  FSFB_LAUNCH(k_project_sample_select_hwc, grid, 256, 0, (cudaStream_t)stream, xyz, n, xyz_stride, lidar2img, cams,
              (const uint4*)mask_hwc16, classes, H, W, ids_sel, cam_sel, fg, overlap, anno, anno_rows, anno_cols, anno_col, scores);
  return FSFB_OK;
}

}  // extern "C"

// ---- a11 frustum point expansion (FSF.extract_fg_pts + double_overlap_pts + get_sir_coors) ----------
// Reference: projects/mmdet3d_plugin/models/detectors/FSF.py:260-308, 357-365.  A foreground point seen
// by k >= 1 (camera, class) masks yields k rows: row `fg position` carries its largest object id; the
// remaining k-1 rows are appended after all n_fg first rows, grouped by k ascending, then by rank
// (2nd, 3rd ... largest id), then by foreground order — the order the reference's cat/repeat/topk
// sequence produces.  The ids are recovered by re-sampling the planes for foreground points only.
namespace fsfb {

constexpr int kMaxOverlap = 16;

// sorted position j over the CSR of overlap counts (groups k = 0..kMaxOverlap)
template <typename MaskT>
__global__ void __launch_bounds__(256)
    k_frustum_expand(const float* __restrict__ xyz, int64_t stride, const float* __restrict__ lidar2img, int cams,
                     const MaskT* __restrict__ mask, int classes, int H, int W, const int32_t* __restrict__ idx_fg,
                     int64_t n_fg, const int32_t* __restrict__ perm, const int32_t* __restrict__ seg,
                     const int32_t* __restrict__ offsets, const int32_t* __restrict__ batch,
                     int32_t* __restrict__ rows_point, int32_t* __restrict__ sir_coors, int32_t* __restrict__ status) {
  __shared__ CamSet s_cams;
  __shared__ int s_extra_base[kMaxOverlap + 2];
  load_cams(s_cams, lidar2img, cams);
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int k = 0; k <= kMaxOverlap; ++k) {
      s_extra_base[k] = acc;
      const int cnt = offsets[k + 1] - offsets[k];
      if (k >= 2) acc += cnt * (k - 1);
    }
  }
  __syncthreads();
  const int64_t plane = (int64_t)H * W;
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n_fg; j += (int64_t)gridDim.x * blockDim.x) {
    const int k = seg[j];
    const int f = perm[j];          // position among foreground points
    const int p = idx_fg[f];        // position among all points
    const float* q = xyz + (int64_t)p * stride;
    const float x = __ldg(q), y = __ldg(q + 1), z = __ldg(q + 2);
    int ids[kMaxOverlap];
    int cnt = 0;
    for (int cam = 0; cam < cams; ++cam) {
      const int tex = project_texel(s_cams.P[cam], x, y, z, W, H);
      if (tex < 0) continue;
      const MaskT* m0 = mask + (int64_t)cam * classes * plane + tex;
      for (int c = 0; c < classes; ++c) {
        const int id = (int)__ldg(m0 + c * plane);
        if (id > 0) {
          // insertion into a descending list (topk order)
          int pos = min(cnt, kMaxOverlap - 1);
          if (cnt >= kMaxOverlap && id <= ids[kMaxOverlap - 1]) continue;
          while (pos > 0 && ids[pos - 1] < id) {
            ids[pos] = ids[pos - 1];
            --pos;
          }
          ids[pos] = id;
          cnt = min(cnt + 1, kMaxOverlap);
        }
      }
    }
    if (cnt != k || k < 1) atomicOr(status, 4);  // inconsistent with the overlap count given (or > kMaxOverlap)
    const int b = batch ? __ldg(batch + p) : 0;
    // first row: foreground order, largest id
    rows_point[f] = p;
    sir_coors[(int64_t)f * 3 + 0] = b;
    sir_coors[(int64_t)f * 3 + 1] = 0;
    sir_coors[(int64_t)f * 3 + 2] = cnt > 0 ? ids[0] : 0;
    const int rank = (int)(j - offsets[k]);
    const int cnt_k = offsets[k + 1] - offsets[k];
    for (int pad = 1; pad < k && pad < cnt; ++pad) {
      const int64_t r = n_fg + s_extra_base[k] + (int64_t)(pad - 1) * cnt_k + rank;
      rows_point[r] = p;
      sir_coors[r * 3 + 0] = b;
      sir_coors[r * 3 + 1] = 0;
      sir_coors[r * 3 + 2] = ids[pad];
    }
  }
}

__global__ void __launch_bounds__(256)
    k_gather_overlap(const uint8_t* __restrict__ overlap, const int32_t* __restrict__ idx_fg, int64_t n_fg,
                     int32_t* __restrict__ ov32) {
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n_fg; j += (int64_t)gridDim.x * blockDim.x)
    ov32[j] = min((int)overlap[idx_fg[j]], kMaxOverlap + 1);  // group kMaxOverlap + 1 = "more than the kernels keep": the host raises
}

}  // namespace fsfb

extern "C" {

int fsfb_gather_overlap(const uint8_t* overlap, const int32_t* idx_fg, int64_t n_fg, int32_t* ov32, void* stream) {
  using namespace fsfb;
  FSFB_CHECK_ARG(n_fg >= 0, "gather_overlap: bad n");
  if (n_fg == 0) return FSFB_OK;
  FSFB_CHECK_ARG(overlap && idx_fg && ov32, "gather_overlap: null pointer");
  const int grid = (int)std::min<int64_t>(ceil_div(n_fg, 256), (int64_t)kNumSMs * 8);
  FSFB_LAUNCH(k_gather_overlap, grid, 256, 0, (cudaStream_t)stream, overlap, idx_fg, n_fg, ov32);
  return FSFB_OK;
}

int fsfb_frustum_expand(const float* xyz, int64_t xyz_stride, const float* lidar2img, int cams, const void* mask,
                        int mask_i32, int classes, int H, int W, const int32_t* idx_fg, int64_t n_fg,
                        const int32_t* perm, const int32_t* seg, const int32_t* offsets, const int32_t* batch_idx,
                        int32_t* rows_point, int32_t* sir_coors, int32_t* status, void* stream) {
  using namespace fsfb;
  int rc = check_common(xyz, n_fg, xyz_stride, lidar2img, cams, mask, classes, H, W, "frustum_expand");
  if (rc != FSFB_OK) return rc;
  if (n_fg == 0) return FSFB_OK;
  FSFB_CHECK_ARG(idx_fg && perm && seg && offsets && rows_point && sir_coors && status, "frustum_expand: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = (int)std::min<int64_t>(ceil_div(n_fg, 256), (int64_t)kNumSMs * 8);
  if (mask_i32) {
    FSFB_LAUNCH(k_frustum_expand<int>, grid, 256, 0, st, xyz, xyz_stride, lidar2img, cams, (const int*)mask, classes, H,
                W, idx_fg, n_fg, perm, seg, offsets, batch_idx, rows_point, sir_coors, status);
  } else {
    FSFB_LAUNCH(k_frustum_expand<unsigned char>, grid, 256, 0, st, xyz, xyz_stride, lidar2img, cams,
                (const unsigned char*)mask, classes, H, W, idx_fg, n_fg, perm, seg, offsets, batch_idx, rows_point,
                sir_coors, status);
  }
  return FSFB_OK;
}

}  // extern "C"
