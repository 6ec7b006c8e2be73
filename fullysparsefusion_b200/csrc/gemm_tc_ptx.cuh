// gemm_tc_ptx.cuh — PTX wrappers (mbarrier, bulk copy, tcgen05) and descriptors shared by the tensor-core
// gather-GEMM kernels (gemm_tc.cu: A and B from shared memory; gemm_ts.cu: A from TMEM).
#pragma once
#include "gemm_common.cuh"

namespace fsfb {

constexpr int kTcRows = 128;           // UMMA M
constexpr int kTcMaxStages = 4;
constexpr int kTcMaxOff = 32;          // koff <= 32 (27 used)
constexpr uint32_t kStageABytes = kTcRows * 128;  // one of hi / lo

// ---- PTX wrappers ----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], kind::tf32
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// SWIZZLE_128B K-major shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
// start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48) | layout=2 [61,64)
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// kind::tf32 instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A=B=tf32, K-major both
__device__ __forceinline__ uint32_t make_idesc_tf32(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kTcRows >> 4) << 24);
}


// shared-memory accessors on 32-bit shared addresses (keeps the accesses in the shared state space)
__device__ __forceinline__ int lds_i32(uint32_t addr) {
  int v;
  asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_i32(uint32_t addr, int v) {
  asm volatile("st.shared.s32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void sts_f4(uint32_t addr, const float4& v) {
  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// predicated (branch-free) read-only loads: a false predicate leaves zeros and issues no request
__device__ __forceinline__ float4 ldg_pred_f4(const float* p, bool pred) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %5, 0;\n\t@p ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];\n\t}"
      : "+f"(v.x), "+f"(v.y), "+f"(v.z), "+f"(v.w)
      : "l"(p), "r"((int)pred));
  return v;
}
__device__ __forceinline__ float ldg_pred_f1(const float* p, bool pred) {
  float v = 0.f;
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %2, 0;\n\t@p ld.global.nc.f32 %0, [%1];\n\t}" : "+f"(v) : "l"(p), "r"((int)pred));
  return v;
}

struct StageCursor {  // walks (active offset k, k-chunk) pairs in issue order
  uint32_t rem;
  int k, kc;
  __device__ __forceinline__ void init(uint32_t mask) {
    rem = mask;
    k = rem ? __ffs(rem) - 1 : 0;
    kc = 0;
  }
  __device__ __forceinline__ void next(int kc_n) {
    if (++kc == kc_n) {
      kc = 0;
      rem &= rem - 1;
      k = rem ? __ffs(rem) - 1 : 0;
    }
  }
};

struct TcParams {
  const float* a;
  int64_t a_rows;
  int cin;
  int64_t a_stride;
  const int32_t* nbr;
  const int32_t* row_order;  // optional: tile row i of the grid handles output row row_order[i] (mask-sorted rows)
  int koff;
  int64_t rows;
  const unsigned char* w_packed;
  GemmShape S;
  Epilogue E;
  float* out;
  int64_t out_stride;
  int stages;
  int out_vec;  // 1: rows of `out` are 16-byte aligned → 128-bit stores
  uint32_t data_bytes;  // stage ring (or epilogue staging, whichever is larger); nbr tile + barriers follow
  // persistent TS kernel (gemm_ts.cu): 128-wide column tiles, offset splits, unit count, split slabs [splits][rows][cpad]
  int n_ct, splits, cpad, n_row_tiles, sched_slot;
  uint32_t sched_base;
  int64_t n_units;
  float* partial;
  uint32_t* timers;  // diagnostics (FSFB_GEMM_TIMERS=1), else null
  // host copies of the per-channel epilogue vectors (cout <= 128): read as constant-bank operands by the epilogue warps
  float hv_bias[128], hv_w[128], hv_h[128];
  int debug;    // FSFB_GEMM_DEBUG bits (profiling experiments only): 1 no A loads, 2 no W copy, 4 no MMA, 8 no A stores
  // shared-memory-operand persistent kernel (gemm_ss.cu): ring depths, column-tile width, accumulator buffers and the byte
  // offsets of the regions behind the A / W rings
  int ss_a_stages, ss_w_stages, ss_tile_w, ss_acc_cols, ss_bufs, ss_stage_stride;
  uint32_t ss_w_slot, ss_off_w, ss_off_nbr, ss_off_stage, ss_off_vec, ss_off_sh;
  int nbr_ro;                 // nbr is permuted into the row order, row stride round_up(rows, 128) (FSFB_NBR_ROW_ORDERED)
  unsigned int* ss_overflow;  // device counter: launches that saw an input outside fp16 range (|a| >= 65504)
};


// ---- shared epilogue -----------------------------------------------------------------------------
// phase 1 (warps 0-3, thread = TMEM lane = tile row): accumulators (main + correction) + bias + norm,
// staged row-major in shared memory at `base` (row stride round_up(n_w,32)+4 floats).
__device__ __forceinline__ void epilogue_phase1(const TcParams& P, uint32_t tmem_d, uint32_t acc_cols, uint32_t base,
                                                int n_w, int nt, bool have_acc, int tid) {
  const int warp = tid >> 5;
  const int n_active = have_acc ? 1 : 0;
  const uint32_t t_row = tmem_d + ((uint32_t)(warp * 32) << 16);
  const Epilogue& E = P.E;
  const int c0 = nt * kGemmNTile;
  const int c_n = min(n_w, P.S.cout - c0);  // real channels in this column tile
  float mean = 0.f, rstd = 1.f;
  float v[32];
  auto tc_ld32 = [&](uint32_t taddr, float(&dst)[32]) {  // main + correction accumulator
    float c2[32];
    fsfb::tc_ld32(taddr, dst);
    fsfb::tc_ld32(taddr + acc_cols, c2);
#pragma unroll
    for (int j = 0; j < 32; ++j) dst[j] += c2[j];
  };
  if (E.norm == FSFB_NORM_LAYERNORM) {  // whole row is in this tile (cout <= 256 enforced on the host)
    float s = 0.f;
    for (int cb = 0; cb < c_n; cb += 32) {
      if (n_active > 0) tc_ld32(t_row + cb, v);
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (cb + j < c_n) s += (n_active > 0 ? v[j] : 0.f) + (E.bias ? __ldg(E.bias + cb + j) : 0.f);
    }
    mean = s / (float)c_n;
    float q = 0.f;
    for (int cb = 0; cb < c_n; cb += 32) {
      if (n_active > 0) tc_ld32(t_row + cb, v);
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (cb + j < c_n) {
          const float d = (n_active > 0 ? v[j] : 0.f) + (E.bias ? __ldg(E.bias + cb + j) : 0.f) - mean;
          q += d * d;
        }
    }
    rstd = 1.f / sqrtf(q / (float)c_n + E.eps);
  }
  // phase 1 (thread = row): bias + norm, staged row-major in the (now idle) stage buffers
  for (int cb = 0; cb < c_n; cb += 32) {
    if (n_active > 0) {
      tc_ld32(t_row + cb, v);  // warp-collective: executed by all lanes, valid row or not
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = 0.f;
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const int c = c0 + cb + j;
      if (cb + j < c_n) {
        float x = v[j] + (E.bias ? __ldg(E.bias + c) : 0.f);
        if (E.norm == FSFB_NORM_LAYERNORM) {
          x = (x - mean) * rstd * __ldg(E.norm_w + c) + __ldg(E.norm_b + c);
        } else if (E.norm == FSFB_NORM_AFFINE) {
          x = fmaf(x, __ldg(E.norm_w + c), __ldg(E.norm_b + c));
        }
        v[j] = x;
      }
    }
    const uint32_t srow = base + (uint32_t)tid * (uint32_t)(((n_w + 31) & ~31) + 4) * 4u + (uint32_t)cb * 4u;
#pragma unroll
    for (int j = 0; j < 32; j += 4) sts_f4(srow + j * 4, make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
  }
}

// phase 2 (n_warps warps, lanes along channels): residual + activation + coalesced global stores.
__device__ __forceinline__ void epilogue_phase2(const TcParams& P, uint32_t base, int n_w, int nt, int64_t row0, int warp,
                                                int lane, int n_warps) {
  const Epilogue& E = P.E;
  const int c0 = nt * kGemmNTile;
  const int c_n = min(n_w, P.S.cout - c0);
  const int act = E.act & 0xff;
  const bool post = (E.act & FSFB_RESIDUAL_POST) != 0;
  const bool res_vec = E.residual && ((uintptr_t)E.residual % 16 == 0) && (E.residual_stride % 4 == 0);
  for (int rl = warp; rl < kTcRows; rl += n_warps) {
    if (row0 + rl >= P.rows || (P.debug & 16)) break;
    const int64_t r = P.row_order ? (int64_t)__ldg(P.row_order + row0 + rl) : row0 + rl;
    const uint32_t srow = base + (uint32_t)rl * (uint32_t)(((n_w + 31) & ~31) + 4) * 4u;
    float* o = P.out + r * P.out_stride + c0;
    const float* res = E.residual ? E.residual + r * E.residual_stride + c0 : nullptr;
    if (P.out_vec && (c_n & 3) == 0 && (!E.residual || res_vec)) {
      for (int c = lane * 4; c < c_n; c += 128) {
        float4 x;
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w) : "r"(srow + c * 4));
        float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
        if (res) q = __ldg(reinterpret_cast<const float4*>(res + c));
        if (post) {
          x.x = apply_act(x.x, act) + q.x; x.y = apply_act(x.y, act) + q.y; x.z = apply_act(x.z, act) + q.z; x.w = apply_act(x.w, act) + q.w;
        } else {
          x.x = apply_act(x.x + q.x, act); x.y = apply_act(x.y + q.y, act); x.z = apply_act(x.z + q.z, act); x.w = apply_act(x.w + q.w, act);
        }
        *reinterpret_cast<float4*>(o + c) = x;
      }
    } else {
      for (int c = lane; c < c_n; c += 32) {
        float x;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"(srow + c * 4));
        const float q = res ? __ldg(res + c) : 0.f;
        o[c] = post ? apply_act(x, act) + q : apply_act(x + q, act);
      }
    }
  }
}

// gemm_ss.cu: returns 1 when the shape is left to the kernels below
int launch_gather_gemm_ss(TcParams& P, bool a_vec, bool a_split, float* workspace, size_t workspace_bytes, int splits,
                          const float* host_bias, const float* host_norm_w, const float* host_norm_b, cudaStream_t st);
// gemm_ts.cu
int launch_splitk_epilogue(const TcParams& P, cudaStream_t st);
// gemm_lin.cu: dense Linear layers over many rows, one CTA per 128-row tile (0 = launched, 1 = shape not served there)
int launch_linear_ss(TcParams& P, bool a_vec, int ksplits, float* workspace, size_t workspace_bytes, cudaStream_t st);
int ss_overflow_counter(unsigned int** out);
int ss_timers_buffer(uint32_t** out, cudaStream_t st);
int launch_gather_gemm_ts(TcParams& P, bool a_vec, float* workspace, size_t workspace_bytes, int splits, const float* host_bias,
                          const float* host_norm_w, const float* host_norm_b, cudaStream_t st);

}  // namespace fsfb
