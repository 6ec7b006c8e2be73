// voxelize.cu — dynamic voxelization (a1).
// Reference semantics: mmdet3d.ops.Voxelization(max_num_points=-1) as used at
// projects/mmdet3d_plugin/models/detectors/single_stage_fsd.py:217-219; in-tree formula
// torch.div(p - min, vs, rounding_mode='floor') (:270, :444, :591-593).
// HBM-bound: 12 B read + 12 B written per point (algorithmic 24 B/pt).
#include "common.cuh"

namespace fsfb {

struct VoxelParams {
  float min_x, min_y, min_z;
  float vs_x, vs_y, vs_z;
  int gx, gy, gz;
};

// IEEE fp32 subtract + divide + floor, no FMA contraction, identical on CPU and GPU.
// mode 0: floor((p - lo) / vs)            — the Voxelization kernel's rule.
// mode 1: torch.div(p - lo, vs, rounding_mode='floor') — ATen's div_floor_floating
//         (Python floor division: exact fmod first), which differs from mode 0 when the
//         fp32 quotient rounds up onto an integer.
__device__ __forceinline__ int voxel_coord(float p, float lo, float vs, int mode) {
  float a = __fsub_rn(p, lo);
  if (mode == 0) return __float2int_rd(__fdiv_rn(a, vs));
  float mod = fmodf(a, vs);  // exact
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

__global__ void __launch_bounds__(256) k_voxelize(const float* __restrict__ pts, int64_t n,
                                                  int64_t stride, VoxelParams P, int mode,
                                                  int order_xyz, int check_range,
                                                  int32_t* __restrict__ coors) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += step) {
    const float* p = pts + i * stride;
    float x = __ldg(p), y = __ldg(p + 1), z = __ldg(p + 2);
    // __float2int_rd saturates for huge inputs, which the range test then rejects; NaN converts
    // to 0 (and mode 1 turns +-inf into NaN through fmod), so non-finite inputs are rejected
    // explicitly.
    int cx = voxel_coord(x, P.min_x, P.vs_x, mode);
    int cy = voxel_coord(y, P.min_y, P.vs_y, mode);
    int cz = voxel_coord(z, P.min_z, P.vs_z, mode);
    bool ok = (fabsf(x) <= 3.402823466e38f) & (fabsf(y) <= 3.402823466e38f) &
              (fabsf(z) <= 3.402823466e38f);
    if (check_range) ok &= (cx >= 0) & (cx < P.gx) & (cy >= 0) & (cy < P.gy) & (cz >= 0) & (cz < P.gz);
    int32_t* o = coors + i * 3;
    o[0] = ok ? (order_xyz ? cx : cz) : -1;
    o[1] = ok ? cy : -1;
    o[2] = ok ? (order_xyz ? cz : cx) : -1;
  }
}

}  // namespace fsfb

extern "C" int fsfb_voxelize(const float* pts, int64_t n, int64_t row_stride,
                             const float* range_min, const float* voxel, const int32_t* grid,
                             int floor_mode, int order_xyz, int check_range, int32_t* coors_zyx,
                             void* stream) {
  using namespace fsfb;
  FSFB_CHECK_ARG(n >= 0 && row_stride >= 3, "voxelize: bad n=%lld stride=%lld", (long long)n,
                 (long long)row_stride);
  FSFB_CHECK_ARG(range_min && voxel && (grid || !check_range), "voxelize: null host parameter");
  FSFB_CHECK_ARG(floor_mode == 0 || floor_mode == 1, "voxelize: floor_mode must be 0 or 1");
  if (n == 0) return FSFB_OK;
  FSFB_CHECK_ARG(pts && coors_zyx, "voxelize: null device pointer");
  VoxelParams P{range_min[0], range_min[1], range_min[2], voxel[0], voxel[1], voxel[2],
                grid ? grid[0] : 0, grid ? grid[1] : 0, grid ? grid[2] : 0};
  FSFB_CHECK_ARG(P.vs_x > 0 && P.vs_y > 0 && P.vs_z > 0, "voxelize: voxel size must be > 0");
  int blocks = (int)std::min<int64_t>(ceil_div(n, 256), (int64_t)kNumSMs * 16);
  FSFB_LAUNCH(k_voxelize, blocks, 256, 0, (cudaStream_t)stream, pts, n, row_stride, P,
              floor_mode, order_xyz, check_range, coors_zyx);
  return FSFB_OK;
}
