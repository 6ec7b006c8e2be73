// gemm_common.cuh — packed-weight layout and the fused epilogue shared by the tensor-core
// gather-GEMM (gemm_tc.cu), its CUDA-core cross-check (gemm_simt.cu) and fsfb_rownorm_act.
#pragma once
#include <cstdlib>

#include "common.cuh"

namespace fsfb {

// ---- packed weight layout (written by fsfb_gemm_prepack, read by the tcgen05 kernel) -------
// w[koff][cout][cin] is cut into column tiles of kGemmNTile output channels (the last one
// narrower, widths rounded up to 16 = UMMA N granularity at M=128) and K chunks of 32 floats
// (= one 128-byte swizzle row of tf32).  Block (nt, k, kc) holds [hi|lo][n_w][32] floats:
// hi = w rounded to nearest tf32, lo = tf32(w - hi).
// Inside a block, row n lives at byte n*128 and its 16-byte chunk j at position j ^ (n & 7)
// (the SWIZZLE_128B K-major canonical layout UMMA shared-memory descriptors expect), so a
// block is moved to shared memory with one linear bulk copy.
constexpr int kGemmNTile = 256;
constexpr int kGemmKChunk = 32;

struct GemmShape {
  int koff, cin, cout;
  __host__ __device__ int kc() const { return (cin + kGemmKChunk - 1) / kGemmKChunk; }
  __host__ __device__ int n_pad() const { return (cout + 15) / 16 * 16; }
  __host__ __device__ int n_tiles() const { return (n_pad() + kGemmNTile - 1) / kGemmNTile; }
  __host__ __device__ int n_w(int nt) const { return min(kGemmNTile, n_pad() - nt * kGemmNTile); }
  // bytes of one (nt,k,kc) block: hi + lo
  __host__ __device__ size_t block_bytes(int nt) const { return (size_t)2 * n_w(nt) * 128; }
  __host__ __device__ size_t tile_base(int nt) const {
    return (size_t)nt * koff * kc() * 2 * kGemmNTile * 128;
  }
  __host__ __device__ size_t block_offset(int nt, int k, int kchunk) const {
    return tile_base(nt) + ((size_t)k * kc() + kchunk) * block_bytes(nt);
  }
  __host__ __device__ size_t total_bytes() const {
    return tile_base(n_tiles() - 1) + (size_t)koff * kc() * block_bytes(n_tiles() - 1);
  }
  // ---- fp16-split copy of the same blocks (experimental, FSFB_GEMM_F16=1): appended after the tf32 blocks ----
  // Block (nt,k,kc) holds n_w rows of 128 bytes: halves [0,32) = fp16(w) of the chunk's 32 inputs, halves [32,64) =
  // fp16((w - hi) * 2048), 16-byte chunks swizzled like the tf32 rows.
  __host__ __device__ size_t f16_block_bytes(int nt) const { return (size_t)n_w(nt) * 128; }
  __host__ __device__ size_t f16_tile_base(int nt) const { return total_bytes() + (size_t)nt * koff * kc() * kGemmNTile * 128; }
  __host__ __device__ size_t f16_block_offset(int nt, int k, int kchunk) const {
    return f16_tile_base(nt) + ((size_t)k * kc() + kchunk) * f16_block_bytes(nt);
  }
  __host__ __device__ size_t total_bytes_f16() const {
    return f16_tile_base(n_tiles() - 1) + (size_t)koff * kc() * f16_block_bytes(n_tiles() - 1);
  }
};

// byte offset of half j (0..63: 32 hi then 32 scaled lo) of row n inside an fp16-split block
__host__ __device__ inline uint32_t sw128_offset_f16(int n, int j) {
  return (uint32_t)n * 128u + (uint32_t)((((j >> 3) ^ (n & 7)) << 4) | ((j & 7) << 1));
}
constexpr float kF16LoScale = 2048.f;  // residuals are stored times 2^11 (keeps them in fp16's normal range); exact to undo

// fp16-split operands (default; FSFB_GEMM_F16=0 restores the 3xTF32 kernels and the tf32-only packed weights):
// fsfb_gemm_prepack also writes the fp16-split blocks and the persistent gather-GEMMs run kind::f16 MMAs on them (same
// 22-bit split precision as 3xTF32 at twice the tensor rate; inputs must stay below 65504).
inline bool gemm_f16_enabled() {
  static const bool on = [] {
    const char* e = getenv("FSFB_GEMM_F16");
    return !e || atoi(e) != 0;
  }();
  return on;
}

// byte offset of element (n, j) inside a [rows][32 float] SWIZZLE_128B K-major block
__host__ __device__ inline uint32_t sw128_offset(int n, int j) {
  return (uint32_t)n * 128u + (uint32_t)((((j >> 2) ^ (n & 7)) << 4) | ((j & 3) << 2));
}

// Round-to-nearest tf32 (low 13 mantissa bits zero afterwards), so the tensor core's own
// operand truncation is a no-op and the split is unbiased: x = hi + lo + O(2^-23 |x|).
__device__ __forceinline__ float tf32_rn(float x) {
  // round-half-away on the magnitude bits (what cvt.rna.tf32.f32 does for finite inputs) in two integer
  // ops; Inf/NaN inputs are outside the contract of this path
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}
__device__ __forceinline__ float tf32_hi(float x) { return tf32_rn(x); }
__device__ __forceinline__ float tf32_lo(float x, float hi) { return tf32_rn(x - hi); }

// ---- epilogue ------------------------------------------------------------------------------
struct Epilogue {
  const float* bias;      // [cout] or null
  int norm;               // FSFB_NORM_*
  const float* norm_w;    // [cout]
  const float* norm_b;    // [cout]
  float eps;
  const float* residual;  // [rows, residual_stride] or null
  int64_t residual_stride;
  int act;                // FSFB_ACT_* (| FSFB_RESIDUAL_POST: add the residual after the activation)
};

// GELU(x) = x Phi(x) with erf from Abramowitz & Stegun 7.1.26 (|erf error| <= 1.5e-7) on the two special-function-unit
// approximations (rcp, ex2): 16 instructions against ~30 for erff(), which bounded every LayerNorm + GELU epilogue.  Measured
// against the float64 value over [-12, 12]: max abs error 4.7e-7 (torch's own fp32 gelu: 1.2e-6).  The negative branch returns
// 0.5 x p e directly, so the tail keeps its relative accuracy instead of cancelling in 1 - erf.
__device__ __forceinline__ float gelu_as(float x) {
  const float z = x * 0.70710678118654752440f, az = fabsf(z);
  float t, e;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, az, 1.f)));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  p *= t;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(az * az * -1.4426950408889634f));
  const float h = 0.5f * x, pe = p * e;
  return z >= 0.f ? fmaf(-h, pe, x) : h * pe;
}

__device__ __forceinline__ float apply_act(float x, int act) {
  if (act == FSFB_ACT_RELU) return fmaxf(x, 0.f);
  if (act == FSFB_ACT_GELU) return gelu_as(x);
  return x;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// One warp finishes one row: x (pre-bias accumulators, row-major in global or anywhere
// addressable) → out.  x and out may alias.
__device__ __forceinline__ void warp_row_epilogue(const float* x, int c, const Epilogue& E,
                                                  int64_t row, float* out) {
  const int lane = lane_id();
  float mean = 0.f, rstd = 1.f;
  if (E.norm == FSFB_NORM_LAYERNORM) {
    float s = 0.f;
    for (int j = lane; j < c; j += 32) s += x[j] + (E.bias ? __ldg(E.bias + j) : 0.f);
    mean = warp_sum(s) / (float)c;
    float q = 0.f;
    for (int j = lane; j < c; j += 32) {
      const float d = x[j] + (E.bias ? __ldg(E.bias + j) : 0.f) - mean;
      q += d * d;
    }
    rstd = 1.f / sqrtf(warp_sum(q) / (float)c + E.eps);
  }
  for (int j = lane; j < c; j += 32) {
    float v = x[j] + (E.bias ? __ldg(E.bias + j) : 0.f);
    if (E.norm == FSFB_NORM_LAYERNORM) {
      v = (v - mean) * rstd * __ldg(E.norm_w + j) + __ldg(E.norm_b + j);
    } else if (E.norm == FSFB_NORM_AFFINE) {
      v = fmaf(v, __ldg(E.norm_w + j), __ldg(E.norm_b + j));
    }
    const int act = E.act & 0xff;
    const bool post = (E.act & FSFB_RESIDUAL_POST) != 0;
    const float rsd = E.residual ? __ldg(E.residual + row * E.residual_stride + j) : 0.f;
    out[j] = post ? apply_act(v, act) + rsd : apply_act(v + rsd, act);
  }
}

inline int check_epilogue(int cout, const float* bias, int norm, const float* norm_w,
                          const float* norm_b, int act, const char* who) {
  (void)bias;
  FSFB_CHECK_ARG(norm == FSFB_NORM_NONE || norm == FSFB_NORM_LAYERNORM || norm == FSFB_NORM_AFFINE,
                 "%s: bad norm %d", who, norm);
  FSFB_CHECK_ARG(norm == FSFB_NORM_NONE || (norm_w && norm_b), "%s: norm needs norm_w and norm_b", who);
  FSFB_CHECK_ARG((act & ~FSFB_RESIDUAL_POST) == FSFB_ACT_NONE || (act & ~FSFB_RESIDUAL_POST) == FSFB_ACT_RELU ||
                     (act & ~FSFB_RESIDUAL_POST) == FSFB_ACT_GELU,
                 "%s: bad act %d", who, act);
  FSFB_CHECK_ARG(cout >= 1, "%s: cout must be >= 1", who);
  return FSFB_OK;
}

}  // namespace fsfb
