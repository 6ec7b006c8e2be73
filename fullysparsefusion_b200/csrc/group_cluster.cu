// group_cluster.cu — the per-class-group half of ClusterAssigner, all groups in one pass.
//
// Reference: SingleStageFSD.group_sample's per-group selection (models/detectors/single_stage_fsd.py:822-842: score >
// threshold, "at least one point" fallback :833-835) and ClusterAssigner.forward_single_class (:936-982: BEV voxel of
// the voted centre with the group's voxel size, drop voxels with < min_points, `valid_mask = ~valid_mask` when nothing
// survives :953-955, voxel-mean centres, connected components with the group's distance), which the reference runs
// as a Python loop over the six class groups.  That loop costs ~150 launches and ~30 output-size reads per frame; the
// same arithmetic is done here on ONE list ordered (group, voxel row): the group id rides along as the batch column of
// the ranking and of the CCL, so the only host reads left are four list lengths.  Per-group results are identical to
// the loop (same candidates, same order inside a group, cluster ids renumbered from 0 per group).
#include "common.cuh"

namespace fsfb {

constexpr int kGcMaxGroups = 8;

struct GcThresholds {
  float thr[kGcMaxGroups];
};
struct GcVoxel {
  float lo[3];
  float vs[kGcMaxGroups][3];
};

// flags[g][v] = score[v][g] > thr[g]; counts[g] = number of set flags
__global__ void __launch_bounds__(256)
    k_gc_flags(const float* __restrict__ score, int64_t n, int64_t stride, int G, GcThresholds T, uint8_t* __restrict__ flags,
               int32_t* __restrict__ counts) {
  __shared__ int s_cnt[kGcMaxGroups];
  if (threadIdx.x < kGcMaxGroups) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    for (int g = 0; g < G; ++g) {
      const bool f = __ldg(score + i * stride + g) > T.thr[g];
      flags[(int64_t)g * n + i] = f ? 1 : 0;
      if (f) atomicAdd(&s_cnt[g], 1);
    }
  }
  __syncthreads();
  if (threadIdx.x < G && s_cnt[threadIdx.x]) atomicAdd(counts + threadIdx.x, s_cnt[threadIdx.x]);
}

// a group without any candidate keeps row 0 (single_stage_fsd.py:833-835)
__global__ void k_gc_fix_empty(uint8_t* __restrict__ flags, int64_t n, int G, const int32_t* __restrict__ counts) {
  const int g = threadIdx.x;
  if (g < G && counts[g] == 0 && n > 0) flags[(int64_t)g * n] = 1;
}

// flat = g * n + v  →  grp, vox, row of the [n, G, 3] centre table
__global__ void __launch_bounds__(256)
    k_gc_split(const int32_t* __restrict__ flat, int64_t t, int64_t n, int G, int32_t* __restrict__ grp,
               int32_t* __restrict__ vox, int32_t* __restrict__ cidx) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < t; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t f = flat[i];
    const int g = (int)(f / n);
    const int v = (int)(f - (int64_t)g * n);
    grp[i] = g;
    vox[i] = v;
    cidx[i] = v * G + g;
  }
}

// torch.div(c - lo, vs_g, rounding_mode='floor') per row with the row's group voxel size (:946-950); rows4 = (g, x, y, z)
__device__ __forceinline__ int gc_floor_div(float p, float lo, float vs) {
  const float a = __fsub_rn(p, lo);
  const float mod = fmodf(a, vs);
  float div = __fdiv_rn(__fsub_rn(a, mod), vs);
  if ((mod != 0.f) && ((vs < 0.f) != (mod < 0.f))) div = __fsub_rn(div, 1.f);
  float fl;
  if (div != 0.f) {
    fl = floorf(div);
    if (__fsub_rn(div, fl) > 0.5f) fl = __fadd_rn(fl, 1.f);
  } else {
    fl = 0.f;
  }
  return __float2int_rd(fl);
}
__global__ void __launch_bounds__(256)
    k_gc_voxelize(const float* __restrict__ ctr, int64_t t, const int32_t* __restrict__ grp, GcVoxel V,
                  int32_t* __restrict__ rows4) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < t; i += (int64_t)gridDim.x * blockDim.x) {
    const int g = grp[i];
    int4 r;
    r.x = g;
    r.y = gc_floor_div(__ldg(ctr + 3 * i), V.lo[0], V.vs[g][0]);
    r.z = gc_floor_div(__ldg(ctr + 3 * i + 1), V.lo[1], V.vs[g][1]);
    r.w = gc_floor_div(__ldg(ctr + 3 * i + 2), V.lo[2], V.vs[g][2]);
    reinterpret_cast<int4*>(rows4)[i] = r;
  }
}

// keep[i] = counts[inv[i]] >= min_points; kept[g] = survivors of group g
__global__ void __launch_bounds__(256)
    k_gc_keep(const int32_t* __restrict__ counts, const int32_t* __restrict__ inv, const int32_t* __restrict__ grp, int64_t t,
              int min_points, uint8_t* __restrict__ keep, int32_t* __restrict__ kept) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < t; i += (int64_t)gridDim.x * blockDim.x) {
    const bool k = counts[inv[i]] >= min_points;
    keep[i] = k ? 1 : 0;
    if (k) atomicAdd(kept + grp[i], 1);
  }
}
// a group in which nothing survives keeps everything (`valid_mask = ~valid_mask`, :953-955)
__global__ void __launch_bounds__(256)
    k_gc_keep_all(const int32_t* __restrict__ grp, int64_t t, const int32_t* __restrict__ kept, uint8_t* __restrict__ keep) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < t; i += (int64_t)gridDim.x * blockDim.x)
    if (kept[grp[i]] == 0) keep[i] = 1;
}

// cluster ids restart at 0 in every group: components are numbered by their lowest member and the list is group-major,
// so a group's first component is its smallest label
__global__ void __launch_bounds__(256)
    k_gc_base(const int32_t* __restrict__ labels, const int32_t* __restrict__ batch, int64_t batch_stride, int64_t m,
              int32_t* __restrict__ base) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x)
    atomicMin(base + batch[i * batch_stride], labels[i]);
}
__global__ void __launch_bounds__(256)
    k_gc_relabel(const int32_t* __restrict__ labels, const int32_t* __restrict__ batch, int64_t batch_stride,
                 const int32_t* __restrict__ base, const int32_t* __restrict__ inv, int64_t t, int32_t* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < t; i += (int64_t)gridDim.x * blockDim.x) {
    const int u = inv[i];
    out[i] = labels[u] - base[batch[(int64_t)u * batch_stride]];
  }
}

static int gc_grid(int64_t n) { return (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(n, 256), (int64_t)kNumSMs * 8)); }

}  // namespace fsfb

extern "C" {

int fsfb_group_flags(const float* score, int64_t n, int64_t stride, int n_groups, const float* thresholds, uint8_t* flags,
                     int32_t* counts, void* stream) {
  using namespace fsfb;
  FSFB_CHECK_ARG(n >= 0 && n_groups >= 1 && n_groups <= kGcMaxGroups && stride >= n_groups && thresholds,
                 "group_flags: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  FSFB_CHECK_ARG(counts, "group_flags: null pointer");
  FSFB_CUDA(cudaMemsetAsync(counts, 0, sizeof(int32_t) * n_groups, st));
  if (n == 0) return FSFB_OK;
  FSFB_CHECK_ARG(score && flags, "group_flags: null pointer");
  GcThresholds T;
  for (int g = 0; g < n_groups; ++g) T.thr[g] = thresholds[g];
  FSFB_LAUNCH(k_gc_flags, gc_grid(n), 256, 0, st, score, n, stride, n_groups, T, flags, counts);
  FSFB_LAUNCH(k_gc_fix_empty, 1, 32, 0, st, flags, n, n_groups, counts);
  return FSFB_OK;
}

int fsfb_group_split(const int32_t* flat, int64_t t, int64_t n, int n_groups, int32_t* grp, int32_t* vox, int32_t* cidx,
                     void* stream) {
  using namespace fsfb;
  FSFB_CHECK_ARG(t >= 0 && n >= 1 && n_groups >= 1 && n * n_groups < (1ll << 31), "group_split: bad argument");
  if (t == 0) return FSFB_OK;
  FSFB_CHECK_ARG(flat && grp && vox && cidx, "group_split: null pointer");
  FSFB_LAUNCH(k_gc_split, gc_grid(t), 256, 0, (cudaStream_t)stream, flat, t, n, n_groups, grp, vox, cidx);
  return FSFB_OK;
}

int fsfb_group_voxelize(const float* centers, int64_t t, const int32_t* grp, const float* range_min, const float* voxel_sizes,
                        int n_groups, int32_t* rows4, void* stream) {
  using namespace fsfb;
  FSFB_CHECK_ARG(t >= 0 && n_groups >= 1 && n_groups <= kGcMaxGroups && range_min && voxel_sizes, "group_voxelize: bad argument");
  if (t == 0) return FSFB_OK;
  FSFB_CHECK_ARG(centers && grp && rows4 && ((uintptr_t)rows4 & 15) == 0, "group_voxelize: null or unaligned pointer");
  GcVoxel V;
  for (int d = 0; d < 3; ++d) V.lo[d] = range_min[d];
  for (int g = 0; g < n_groups; ++g)
    for (int d = 0; d < 3; ++d) {
      V.vs[g][d] = voxel_sizes[3 * g + d];
      FSFB_CHECK_ARG(V.vs[g][d] > 0, "group_voxelize: voxel size must be > 0");
    }
  FSFB_LAUNCH(k_gc_voxelize, gc_grid(t), 256, 0, (cudaStream_t)stream, centers, t, grp, V, rows4);
  return FSFB_OK;
}

int fsfb_group_keep(const int32_t* counts, const int32_t* inv, const int32_t* grp, int64_t t, int n_groups, int min_points,
                    uint8_t* keep, int32_t* kept_per_group, void* stream) {
  using namespace fsfb;
  FSFB_CHECK_ARG(t >= 0 && n_groups >= 1 && n_groups <= kGcMaxGroups && kept_per_group, "group_keep: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  FSFB_CUDA(cudaMemsetAsync(kept_per_group, 0, sizeof(int32_t) * n_groups, st));
  if (t == 0) return FSFB_OK;
  FSFB_CHECK_ARG(counts && inv && grp && keep, "group_keep: null pointer");
  FSFB_LAUNCH(k_gc_keep, gc_grid(t), 256, 0, st, counts, inv, grp, t, min_points, keep, kept_per_group);
  FSFB_LAUNCH(k_gc_keep_all, gc_grid(t), 256, 0, st, grp, t, kept_per_group, keep);
  return FSFB_OK;
}

int fsfb_group_relabel(const int32_t* labels, const int32_t* batch, int64_t batch_stride, int64_t m, int n_groups,
                       const int32_t* inv, int64_t t, int32_t* base, int32_t* out, void* stream) {
  using namespace fsfb;
  FSFB_CHECK_ARG(m >= 0 && t >= 0 && n_groups >= 1 && n_groups <= kGcMaxGroups && base && batch_stride >= 1,
                 "group_relabel: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  FSFB_CUDA(cudaMemsetAsync(base, 0x7f, sizeof(int32_t) * n_groups, st));
  if (m == 0 || t == 0) return FSFB_OK;
  FSFB_CHECK_ARG(labels && batch && inv && out, "group_relabel: null pointer");
  FSFB_LAUNCH(k_gc_base, gc_grid(m), 256, 0, st, labels, batch, batch_stride, m, base);
  FSFB_LAUNCH(k_gc_relabel, gc_grid(t), 256, 0, st, labels, batch, batch_stride, base, inv, t, out);
  return FSFB_OK;
}

}  // extern "C"
