// fsd_ops.cu — small per-point / per-instance steps of the query-generation stages.
//   group_sample      SingleStageFSD.group_sample (models/detectors/single_stage_fsd.py:802-865) +
//                     FSF.get_point_fg_weights (models/detectors/FSF.py:345-355): softmax, per-group score
//                     sums, max-logit offset weights, vote-shifted centres — one pass over [n, C+1] logits.
//   weighted_xyz      get_cluster_delta_weighted input (FSF.py:313-318): (xyz*w, w), w = clamp(w, 1e-5)
//   cluster_delta     centre = mean[:, :3] / mean[:, 3] and f_cluster = xyz - centre[inv] (FSF.py:324-329);
//                     also serves extract_feat's f_cluster (single_stage_fsd.py:460-462) with unit weights
//   encode_preds_2d   get_single_cls_preds_2d + encode_preds_2d (FSF.py:449-504)
// HBM-bound elementwise passes; fp32 without FMA contraction where the reference composes ATen ops.
#include "common.cuh"

namespace fsfb {

constexpr int kGsMaxClasses = 32;
constexpr int kGsMaxGroups = 8;
constexpr int kGsMaxPerGroup = 12;  // AV2 groups hold up to 9 classes (FSF_AV2_config.py:40-45)

struct GroupSpec {
  int n_groups;
  int len[kGsMaxGroups];
  int cls[kGsMaxGroups][kGsMaxPerGroup];
};

__global__ void __launch_bounds__(256)
    k_group_sample(const float* __restrict__ logits, int64_t n, int c1 /* classes + background */,
                   const float* __restrict__ xyz, int64_t xyz_stride, const float* __restrict__ offsets /* [n,c1*3] */,
                   GroupSpec G, float* __restrict__ fg_weight, float* __restrict__ group_score,
                   float* __restrict__ group_center) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float l[kGsMaxClasses];
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < kGsMaxClasses; ++j)
      if (j < c1) {
        l[j] = __ldg(logits + i * c1 + j);
        mx = fmaxf(mx, l[j]);
      }
    float e[kGsMaxClasses], sum = 0.f;
#pragma unroll
    for (int j = 0; j < kGsMaxClasses; ++j)
      if (j < c1) {
        e[j] = expf(__fsub_rn(l[j], mx));
        sum = __fadd_rn(sum, e[j]);
      }
    if (fg_weight) fg_weight[i] = __fsub_rn(1.f, __fdiv_rn(e[c1 - 1], sum));
    if (!group_score) continue;
    const float px = xyz ? __ldg(xyz + i * xyz_stride) : 0.f, py = xyz ? __ldg(xyz + i * xyz_stride + 1) : 0.f,
                pz = xyz ? __ldg(xyz + i * xyz_stride + 2) : 0.f;
    for (int g = 0; g < G.n_groups; ++g) {
      float s = 0.f, gmax = -INFINITY;
      for (int t = 0; t < G.len[g]; ++t) {
        const int c = G.cls[g][t];
        s = __fadd_rn(s, __fdiv_rn(e[c], sum));
        gmax = fmaxf(gmax, l[c]);
      }
      group_score[i * G.n_groups + g] = s;
      if (group_center) {
        // get_offset_weight('max'): indicator(|logit - max| < 1e-6) / count; centre = xyz + sum_j w_j * offset_j
        float wsum = 0.f;
        for (int t = 0; t < G.len[g]; ++t) wsum += (fabsf(__fsub_rn(l[G.cls[g][t]], gmax)) < 1e-6f) ? 1.f : 0.f;
        float ox = 0.f, oy = 0.f, oz = 0.f;
        for (int t = 0; t < G.len[g]; ++t) {
          const int c = G.cls[g][t];
          const float w = __fdiv_rn((fabsf(__fsub_rn(l[c], gmax)) < 1e-6f) ? 1.f : 0.f, wsum);
          const float* o = offsets + i * (int64_t)c1 * 3 + c * 3;
          ox = __fadd_rn(ox, __fmul_rn(__ldg(o), w));
          oy = __fadd_rn(oy, __fmul_rn(__ldg(o + 1), w));
          oz = __fadd_rn(oz, __fmul_rn(__ldg(o + 2), w));
        }
        float* cc = group_center + (i * G.n_groups + g) * 3;
        cc[0] = __fadd_rn(px, ox);
        cc[1] = __fadd_rn(py, oy);
        cc[2] = __fadd_rn(pz, oz);
      }
    }
  }
}

// out[r] = (x*w, y*w, z*w, w), w = max(weight[src], 1e-5); src = rows ? rows[r] : r
__global__ void __launch_bounds__(256)
    k_weighted_xyz(const float* __restrict__ xyz, int64_t stride, const float* __restrict__ weight,
                   const int32_t* __restrict__ rows, int64_t n_rows, float* __restrict__ out) {
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = rows ? rows[r] : r;
    const float w = fmaxf(__ldg(weight + p), 1e-5f);
    const float* q = xyz + p * stride;
    reinterpret_cast<float4*>(out)[r] = make_float4(__fmul_rn(__ldg(q), w), __fmul_rn(__ldg(q + 1), w), __fmul_rn(__ldg(q + 2), w), w);
  }
}

// centre[k] = mean4[k, :3] / mean4[k, 3] (weighted) or mean[k, :3] (plain, c == 3);
// f_cluster[r] = xyz[src] - centre[inv[r]]
__global__ void __launch_bounds__(256)
    k_cluster_center(const float* __restrict__ mean, int64_t k, int c, float* __restrict__ center) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < k; i += (int64_t)gridDim.x * blockDim.x) {
    const float* m = mean + i * c;
    const float d = c == 4 ? __ldg(m + 3) : 1.f;
    center[i * 3 + 0] = c == 4 ? __fdiv_rn(__ldg(m), d) : __ldg(m);
    center[i * 3 + 1] = c == 4 ? __fdiv_rn(__ldg(m + 1), d) : __ldg(m + 1);
    center[i * 3 + 2] = c == 4 ? __fdiv_rn(__ldg(m + 2), d) : __ldg(m + 2);
  }
}

__global__ void __launch_bounds__(256)
    k_cluster_delta(const float* __restrict__ xyz, int64_t stride, const int32_t* __restrict__ rows, int64_t n_rows,
                    const float* __restrict__ center, const int32_t* __restrict__ inv, float* __restrict__ f_cluster) {
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = rows ? rows[r] : r;
    const float* q = xyz + p * stride;
    const float* cc = center + (int64_t)inv[r] * 3;
    f_cluster[r * 3 + 0] = __fsub_rn(__ldg(q), __ldg(cc));
    f_cluster[r * 3 + 1] = __fsub_rn(__ldg(q + 1), __ldg(cc + 1));
    f_cluster[r * 3 + 2] = __fsub_rn(__ldg(q + 2), __ldg(cc + 2));
  }
}

// preds[k] = id >= 1 ? anno[id-1] : (0,..,category = num_classes,..); feat = [bbox/wh, score, onehot(category, num_classes+1)]
__global__ void __launch_bounds__(128)
    k_encode_preds_2d(const float* __restrict__ anno, int anno_rows, int anno_cols, const int32_t* __restrict__ coors,
                      int coor_stride, int coor_col, int64_t k, float img_w, float img_h, int num_classes,
                      float* __restrict__ preds, float* __restrict__ feat) {
  const int fc = 5 + num_classes + 1;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < k; i += (int64_t)gridDim.x * blockDim.x) {
    const int id = coors[i * coor_stride + coor_col] - 1;
    float row[16];
    for (int j = 0; j < anno_cols && j < 16; ++j) row[j] = (id >= 0 && id < anno_rows) ? __ldg(anno + (int64_t)id * anno_cols + j) : 0.f;
    if (id < 0) row[5] = (float)num_classes;
    for (int j = 0; j < anno_cols && j < 16; ++j) preds[i * anno_cols + j] = row[j];
    float* f = feat + i * fc;
    f[0] = __fdiv_rn(row[0], img_w);
    f[1] = __fdiv_rn(row[1], img_h);
    f[2] = __fdiv_rn(row[2], img_w);
    f[3] = __fdiv_rn(row[3], img_h);
    f[4] = row[4];
    const int cat = (int)row[5];
    for (int j = 0; j <= num_classes; ++j) f[5 + j] = (j == cat) ? 1.f : 0.f;
  }
}

static int grid1(int64_t n) { return (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(n, 256), (int64_t)kNumSMs * 8)); }

}  // namespace fsfb

extern "C" {

int fsfb_group_sample(const float* logits, int64_t n, int num_classes_with_bg, const float* xyz, int64_t xyz_stride,
                      const float* offsets, const int32_t* group_lens, const int32_t* group_classes, int n_groups,
                      float* fg_weight, float* group_score, float* group_center, void* stream) {
  using namespace fsfb;
  FSFB_CHECK_ARG(n >= 0 && num_classes_with_bg >= 2 && num_classes_with_bg <= kGsMaxClasses, "group_sample: bad class count");
  FSFB_CHECK_ARG(n_groups >= 0 && n_groups <= kGsMaxGroups, "group_sample: at most %d groups", kGsMaxGroups);
  GroupSpec G;
  G.n_groups = n_groups;
  int pos = 0;
  for (int g = 0; g < n_groups; ++g) {
    FSFB_CHECK_ARG(group_lens && group_classes && group_lens[g] >= 1 && group_lens[g] <= kGsMaxPerGroup,
                   "group_sample: group size must be 1..%d", kGsMaxPerGroup);
    G.len[g] = group_lens[g];
    for (int t = 0; t < group_lens[g]; ++t) {
      const int c = group_classes[pos++];
      FSFB_CHECK_ARG(c >= 0 && c < num_classes_with_bg - 1, "group_sample: class index out of range");
      G.cls[g][t] = c;
    }
  }
  if (n == 0) return FSFB_OK;
  FSFB_CHECK_ARG(logits && (fg_weight || group_score), "group_sample: null pointer");
  FSFB_CHECK_ARG(!group_center || (xyz && offsets && group_score), "group_sample: centres need xyz, offsets and scores");
  FSFB_LAUNCH(k_group_sample, grid1(n), 256, 0, (cudaStream_t)stream, logits, n, num_classes_with_bg, xyz, xyz_stride,
              offsets, G, fg_weight, n_groups ? group_score : (float*)nullptr, group_center);
  return FSFB_OK;
}

int fsfb_weighted_xyz(const float* xyz, int64_t xyz_stride, const float* weight, const int32_t* rows, int64_t n_rows,
                      float* out4, void* stream) {
  using namespace fsfb;
  FSFB_CHECK_ARG(n_rows >= 0 && xyz_stride >= 3, "weighted_xyz: bad argument");
  if (n_rows == 0) return FSFB_OK;
  FSFB_CHECK_ARG(xyz && weight && out4 && ((uintptr_t)out4 & 15) == 0, "weighted_xyz: null/unaligned pointer");
  FSFB_LAUNCH(k_weighted_xyz, grid1(n_rows), 256, 0, (cudaStream_t)stream, xyz, xyz_stride, weight, rows, n_rows, out4);
  return FSFB_OK;
}

int fsfb_cluster_delta(const float* xyz, int64_t xyz_stride, const int32_t* rows, int64_t n_rows, const float* mean,
                       int64_t k, int mean_cols, const int32_t* inv, float* center, float* f_cluster, void* stream) {
  using namespace fsfb;
  FSFB_CHECK_ARG(n_rows >= 0 && k >= 0 && (mean_cols == 3 || mean_cols == 4) && xyz_stride >= 3, "cluster_delta: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (k > 0) {
    FSFB_CHECK_ARG(mean && center, "cluster_delta: null pointer");
    FSFB_LAUNCH(k_cluster_center, grid1(k), 256, 0, st, mean, k, mean_cols, center);
  }
  if (n_rows > 0 && f_cluster) {
    FSFB_CHECK_ARG(xyz && inv && center, "cluster_delta: null pointer");
    FSFB_LAUNCH(k_cluster_delta, grid1(n_rows), 256, 0, st, xyz, xyz_stride, rows, n_rows, center, inv, f_cluster);
  }
  return FSFB_OK;
}

int fsfb_encode_preds_2d(const float* anno, int anno_rows, int anno_cols, const int32_t* obj_coors, int coor_stride,
                         int coor_col, int64_t k, float img_w, float img_h, int num_classes, float* preds_2d,
                         float* feat, void* stream) {
  using namespace fsfb;
  FSFB_CHECK_ARG(k >= 0 && anno_cols >= 6 && anno_cols <= 16 && num_classes >= 1 && coor_col >= 0 && coor_col < coor_stride,
                 "encode_preds_2d: bad argument");
  if (k == 0) return FSFB_OK;
  FSFB_CHECK_ARG(anno && obj_coors && preds_2d && feat, "encode_preds_2d: null pointer");
  FSFB_LAUNCH(k_encode_preds_2d, (int)ceil_div(k, 128), 128, 0, (cudaStream_t)stream, anno, anno_rows, anno_cols, obj_coors,
              coor_stride, coor_col, k, img_w, img_h, num_classes, preds_2d, feat);
  return FSFB_OK;
}

}  // extern "C"
