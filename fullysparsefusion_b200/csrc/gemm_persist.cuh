// gemm_persist.cuh — device helpers shared by the persistent gather-GEMM kernels (gemm_ts.cu: A operand through tensor
// memory; gemm_ss.cu: A operand through a shared-memory ring): tcgen05 wrappers, the work-unit decoding and the
// thread-per-row epilogue over 32-column TMEM blocks.
#pragma once
#include <cuda_fp16.h>

#include "gemm_tc_ptx.cuh"

namespace fsfb {

__device__ __forceinline__ void tc_mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// kind::f16 twin (fp16-split operands, FSFB_GEMM_F16=1): A = 128 x 16 halves in 8 TMEM columns, B = 16 halves per row
__device__ __forceinline__ void tc_mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void sts_u4(uint32_t addr, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], kind::f16 (fp16 operands, fp32 accumulate)
__device__ __forceinline__ void tc_mma_f16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ float4 ldg_pred_f4_na(const float* p, bool pred) {  // read-only path, no L1 allocation
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %5, 0;\n\t@p ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];\n\t}"
      : "+f"(v.x), "+f"(v.y), "+f"(v.z), "+f"(v.w)
      : "l"(p), "r"((int)pred));
  return v;
}

// kind::f16 instruction descriptor: D = f32, A = B = fp16 (format 0), K-major both
__device__ __forceinline__ uint32_t make_idesc_f16(int n) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kTcRows >> 4) << 24);
}

__device__ __forceinline__ void tc_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
      "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}

__device__ __forceinline__ void tc_ld8_nowait(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}

__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  float4 x;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w) : "r"(addr));
  return x;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}

// Producer-side stage walk: a group steps kTsAStages stages at a time through (active offset, K chunk) pairs.
struct TsCursor {
  uint32_t rem;
  int kc;
  __device__ __forceinline__ void init(uint32_t mask, int skip, int kc_n) {
    rem = mask;
    kc = 0;
    step(skip, kc_n);
  }
  __device__ __forceinline__ void step(int n, int kc_n) {
    kc += n;
    while (kc >= kc_n && rem) {
      kc -= kc_n;
      rem &= rem - 1;
    }
  }
  __device__ __forceinline__ int k() const { return __ffs(rem) - 1; }
};

// The unit sequence of this CTA and what every role needs to know about a unit.
struct TsUnit {
  int64_t row0;      // first tile row
  int ct;            // 128-wide column tile
  int n_sub;         // its width (multiple of 16)
  uint32_t k_keep;   // offsets of this unit's split
  int sp;            // split index
};

__device__ __forceinline__ bool ts_unit(const TcParams& P, uint32_t u, TsUnit& U) {
  if (u >= (uint32_t)P.n_units) return false;  // n_units < 2^31 (checked on the host)
  uint32_t rt = u, ct = 0, sp = 0;
  if (P.n_ct * P.splits > 1) {
    const uint32_t per_tile = (uint32_t)(P.n_ct * P.splits);
    rt = u / per_tile;
    const uint32_t rem = u - rt * per_tile;
    ct = rem / (uint32_t)P.splits;
    sp = rem - ct * (uint32_t)P.splits;
  }
  rt = (uint32_t)P.n_row_tiles - 1u - rt;  // from the end of the row order: tiles with the most offsets first
  U.ct = (int)ct;
  U.sp = (int)sp;
  U.row0 = (int64_t)rt * kTcRows;
  U.n_sub = min(128, P.S.n_pad() - U.ct * 128);
  U.k_keep = 0xffffffffu;
  if (P.splits > 1) {
    const uint32_t k_lo = ((uint32_t)P.koff * sp) / (uint32_t)P.splits, k_hi = ((uint32_t)P.koff * (sp + 1)) / (uint32_t)P.splits;
    U.k_keep = (k_hi >= 32 ? 0xffffffffu : ((1u << k_hi) - 1u)) & ~((1u << k_lo) - 1u);
  }
  return true;
}

// Epilogue of one unit in the HV form (per-channel vectors in the kernel parameters), thread = TMEM lane = row, 32 columns
// per tcgen05.ld pair.  NORM / ACT / POST are compile-time so that a fully unrolled instance (UNROLL) is a few hundred
// instructions whose vector operands are constant-bank immediates — no load instruction of any kind (measured: the same
// body rolled, with ld.const indexing, was no faster than shared-memory vectors; GELU layers therefore keep that path).  Leaves the finished row in the staging row `my_row`; arrives on `acc_empty_bar` after the last TMEM read.
template <int NORM, int ACT, bool POST, bool UNROLL, bool F16 = false>
__device__ __forceinline__ void ts_epi32(const TcParams& P, uint32_t t_row, uint32_t acc_cols, uint32_t my_row, int c_n, int n_sub,
                                         bool have_acc, bool use_res, bool valid, uint32_t acc_empty_bar, int lane) {
  float v[32];
  auto ld32 = [&](int cb) {  // warp-collective: executed by all lanes, valid row or not
    if (have_acc) {
      float c2[32];
      tc_ld32(t_row + cb, v);
      tc_ld32(t_row + acc_cols + cb, c2);
      if constexpr (F16) {  // the correction products carry the 2^11 of the scaled residuals
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) v[jj] = fmaf(c2[jj], 1.f / kF16LoScale, v[jj]);
      } else {
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) v[jj] += c2[jj];
      }
    } else {
#pragma unroll
      for (int jj = 0; jj < 32; ++jj) v[jj] = 0.f;
    }
  };
  float mean = 0.f, rstd = 1.f;
  if (NORM == FSFB_NORM_LAYERNORM) {  // two-pass row statistics (the whole row is in this tile)
    float sum = 0.f;
#pragma unroll 1
    for (int cb = 0; cb < c_n; cb += 32) {
      ld32(cb);
#pragma unroll
      for (int jj = 0; jj < 32; ++jj)
        if (cb + jj < c_n) sum += v[jj] + P.hv_bias[cb + jj];
    }
    mean = sum / (float)c_n;
    float qq = 0.f;
#pragma unroll 1
    for (int cb = 0; cb < c_n; cb += 32) {
      ld32(cb);
#pragma unroll
      for (int jj = 0; jj < 32; ++jj)
        if (cb + jj < c_n) {
          const float d = v[jj] + P.hv_bias[cb + jj] - mean;
          qq += d * d;
        }
    }
    rstd = 1.f / sqrtf(qq / (float)c_n + P.E.eps);
  }
  auto block = [&](int cb) {
    ld32(cb);
    if (cb + 32 >= n_sub) {  // last TMEM read of this unit: the MMA warps may start the next unit
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty_bar);
    }
#pragma unroll
    for (int jj = 0; jj < 32; jj += 4) {
      float y[4] = {v[jj], v[jj + 1], v[jj + 2], v[jj + 3]};
      if (cb + jj < c_n) {
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        if (use_res && valid) g = lds_f4(my_row + (uint32_t)(cb + jj) * 4u);
        const float gg[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float x = y[e] + P.hv_bias[cb + jj + e];
          if (NORM == FSFB_NORM_LAYERNORM) x = (x - mean) * rstd * P.hv_w[cb + jj + e] + P.hv_h[cb + jj + e];
          else if (NORM == FSFB_NORM_AFFINE) x = fmaf(x, P.hv_w[cb + jj + e], P.hv_h[cb + jj + e]);
          const float a = apply_act(POST ? x : x + gg[e], ACT);
          y[e] = POST ? a + gg[e] : a;
        }
      }
      sts_f4(my_row + (uint32_t)(cb + jj) * 4u, make_float4(y[0], y[1], y[2], y[3]));
    }
  };
  if (UNROLL) {
#pragma unroll
    for (int cbi = 0; cbi < 4; ++cbi)
      if (32 * cbi < n_sub) block(32 * cbi);
  } else {
#pragma unroll 1
    for (int cb = 0; cb < n_sub; cb += 32) block(cb);
  }
}

}  // namespace fsfb
