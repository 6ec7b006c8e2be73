// gemm_ts.cu — gather-GEMM with the A operand in TENSOR MEMORY (tcgen05.mma "TS" form).
//
// Same contract as gemm_tc.cu (out[r] = epi(sum_k a[nbr[k][r]] @ w[k]^T)); used for column tiles of
// <= 128 channels, i.e. every layer of the stock FSF networks except the 256/512-wide U-Net levels.
//
// Why: with both operands in shared memory, one 128x128x8 tf32 MMA reads 8 KB of smem per ~67 clk, i.e.
// the full 128 B/clk of an SM, and the 3xTF32 scheme re-reads A and W for each of its three products —
// the operand fetch then competes with the producers' stores and the weight copies and the tensor pipe
// sat at ~26 % (ncu, profiles/).  Here the gathered rows never touch shared memory:
//   * 16 producer warps own TMEM lanes (warp w: lane quarter w%4, K quarter w/4): each thread gathers 32 B
//     of ITS row with predicated 128-bit loads, splits fp32 → tf32 hi/lo in registers and writes them with
//     tcgen05.st (32x32b.x8) into a 4-stage ring of TMEM columns;
//   * a loader thread streams the pre-packed W blocks into a deep shared-memory ring with UBLKCP;
//   * the MMA thread issues tcgen05.mma [d], [a_tmem], b_desc (A from TMEM, B from smem), 3 per K-step.
// Shared memory then carries only W (32 KB written + 48 KB read per stage instead of 160 KB).
// TMEM columns: [0,acc) main accumulator, [acc,2acc) correction accumulator, then 4 x 64 columns of A
// stages (32 hi + 32 lo).  512 columns are allocated: one CTA per SM.
#include <cstdlib>

#include "gemm_tc_ptx.cuh"

namespace fsfb {

constexpr int kTsProducers = 512;  // 16 warps: warp w owns TMEM lane quarter w%4 and K quarter w/4
constexpr int kTsThreads = 608;    // + two MMA warps (16, 17) + W loader warp (18)
constexpr int kTsAStages = 4;
constexpr int kTsMaxWStages = 6;

struct TsShared {
  uint64_t a_full[kTsAStages];
  uint64_t a_empty[kTsAStages];
  uint64_t w_full[kTsMaxWStages];
  uint64_t w_empty[kTsMaxWStages];
  uint64_t accum;
  uint32_t tmem_base;
  uint32_t off_mask;
};

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

__device__ __forceinline__ void tc_st8(uint32_t taddr, const float (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(__float_as_uint(v[0])),
               "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])),
               "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
               : "memory");
}

__device__ __forceinline__ void tc_st16(uint32_t taddr, const float (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
      "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
      "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
      "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
      : "memory");
}

template <bool AVEC>
__global__ void __launch_bounds__(kTsThreads, 1) k_gather_gemm_ts(const TcParams P) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nt = blockIdx.y;
  const int n_w = P.S.n_w(nt);
  const int64_t row0 = (int64_t)blockIdx.x * kTcRows;
  const uint32_t w_bytes = (uint32_t)P.S.block_bytes(nt);
  const uint32_t base = smem_u32(smem_raw);
  const uint32_t s_nbr = base + P.data_bytes;  // [koff][128] i32
  TsShared* sh = reinterpret_cast<TsShared*>(smem_raw + (size_t)P.data_bytes + (size_t)P.koff * kTcRows * 4);
  const int w_stages = P.stages;

  if (tid == 0) {
    if (base & 1023u) __trap();
    for (int s = 0; s < kTsAStages; ++s) {
      mbar_init(smem_u32(&sh->a_full[s]), kTsProducers / 32 + 1);  // producer warps + the W loader (with its tx bytes)
      mbar_init(smem_u32(&sh->a_empty[s]), 2);                     // both MMA issuers commit
    }
    for (int s = 0; s < w_stages; ++s) {
      mbar_init(smem_u32(&sh->w_full[s]), 1);
      mbar_init(smem_u32(&sh->w_empty[s]), 1);
    }
    mbar_init(smem_u32(&sh->accum), 2);
    sh->off_mask = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  uint32_t acc_cols = 32;
  while ((int)acc_cols < n_w) acc_cols <<= 1;  // <= 128 (host guarantees n_w <= 128)
  constexpr uint32_t tmem_cols = 512;
  if (warp == kTsProducers / 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sh->tmem_base)),
                 "r"(tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  __syncthreads();
  if (tid < kTsProducers) {  // neighbour tile + active-offset mask (as in gemm_tc.cu)
    const int r_l = tid & (kTcRows - 1);
    const int64_t r = row0 + r_l;
    uint32_t my_mask = 0;
    constexpr int kPar = kTsProducers / kTcRows;
    for (int k0 = tid / kTcRows; k0 < P.koff; k0 += 4 * kPar) {
      int32_t src[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int k = k0 + u * kPar;
        src[u] = -1;
        if (k < P.koff && r < P.rows) src[u] = P.nbr ? __ldg(P.nbr + (int64_t)k * P.rows + r) : (int32_t)r;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int k = k0 + u * kPar;
        if (k < P.koff) {
          if (src[u] >= P.a_rows) src[u] = -1;
          sts_i32(s_nbr + (uint32_t)(k * kTcRows + r_l) * 4u, src[u]);
          my_mask |= (src[u] >= 0 ? 1u : 0u) << k;
        }
      }
    }
    my_mask = __reduce_or_sync(0xffffffffu, my_mask);
    if (lane == 0 && my_mask) atomicOr(&sh->off_mask, my_mask);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t off_mask = sh->off_mask;
  const uint32_t tmem_d = sh->tmem_base;
  const uint32_t tmem_a = tmem_d + 2 * acc_cols;  // A stage ring: stage s at +64*s (hi 32 cols | lo 32 cols)
  const int kc_n = P.S.kc();
  const int n_active = __popc(off_mask) * kc_n;

  if (tid < kTsProducers) {
    // ================= A producers: thread = (row, K half) → TMEM =================
    constexpr int kF4 = 32 * kTcRows / kTsProducers / 4;  // float4 loads per thread per stage (2)
    const int q = warp & 3, h = warp >> 2;
    const int row = 32 * q + lane;
    const uint32_t t_lane = (uint32_t)(32 * q) << 16;
    auto load_stage = [&](const StageCursor& c, float4(&v)[kF4]) {
      const int32_t src = lds_i32(s_nbr + (uint32_t)(c.k * kTcRows + row) * 4u);
      const int col0 = c.kc * kGemmKChunk + 4 * kF4 * h;
#pragma unroll
      for (int j = 0; j < kF4; ++j) {
        const int col = col0 + 4 * j;
        const bool ok = src >= 0 && col < P.cin && !(P.debug & 1);
        const float* g = P.a + (int64_t)(ok ? src : 0) * P.a_stride + (ok ? col : 0);
        if (AVEC && col + 4 <= P.cin) {
          v[j] = ldg_pred_f4(g, ok);
        } else {
          v[j].x = ldg_pred_f1(g, ok);
          v[j].y = ldg_pred_f1(g + 1, ok && col + 1 < P.cin);
          v[j].z = ldg_pred_f1(g + 2, ok && col + 2 < P.cin);
          v[j].w = ldg_pred_f1(g + 3, ok && col + 3 < P.cin);
        }
      }
    };
    float4 cur[kF4], n1[kF4], n2[kF4];
    StageCursor c_ld;
    c_ld.init(off_mask);
    if (n_active > 0) load_stage(c_ld, cur);
    c_ld.next(kc_n);
    if (n_active > 1) load_stage(c_ld, n1);
    c_ld.next(kc_n);
    int s = 0;
    uint32_t ph = 0;
    for (int it = 0; it < n_active; ++it) {
      if (it + 2 < n_active) load_stage(c_ld, n2);
      c_ld.next(kc_n);
      // hi = x with the 13 low mantissa bits cleared (what the tensor core keeps of a tf32 operand),
      // lo = x - hi exactly (the MMA truncates it to tf32 itself): 2 ALU ops per element
      float hi[4 * kF4], lo[4 * kF4];
#pragma unroll
      for (int j = 0; j < kF4; ++j) {
        const float x[4] = {cur[j].x, cur[j].y, cur[j].z, cur[j].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          hi[4 * j + e] = __uint_as_float(__float_as_uint(x[e]) & 0xFFFFE000u);
          lo[4 * j + e] = x[e] - hi[4 * j + e];
        }
      }
      if (lane == 0) mbar_wait(smem_u32(&sh->a_empty[s]), ph ^ 1u);
      __syncwarp();
      tc_fence_after();
      const uint32_t ta = tmem_a + t_lane + (uint32_t)(64 * s + 4 * kF4 * h);
      if (!(P.debug & 8)) {
        tc_st8(ta, hi);
        tc_st8(ta + 32, lo);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&sh->a_full[s]));
#pragma unroll
      for (int j = 0; j < kF4; ++j) {
        cur[j] = n1[j];
        n1[j] = n2[j];
      }
      if (++s == kTsAStages) {
        s = 0;
        ph ^= 1u;
      }
    }
  } else if ((warp == kTsProducers / 32 || warp == kTsProducers / 32 + 1) && lane == 0) {
    // ================= two MMA issuers =================
    // warp 16: main accumulator (a_hi * w_hi); warp 17: correction accumulator (a_lo * w_hi + a_hi * w_lo).
    // The two instruction streams touch different TMEM accumulators, so they need no mutual ordering; splitting
    // them doubles the issue rate of the single-thread MMA front end (the limiter at 12 MMAs per 32-wide K chunk).
    const bool is_main = warp == kTsProducers / 32;
    const uint32_t idesc = make_idesc_tf32(n_w);
    const uint32_t d_acc = is_main ? tmem_d : tmem_d + acc_cols;
    StageCursor c;
    c.init(off_mask);
    int s = 0;
    uint32_t ph = 0;
    for (int it = 0; it < n_active; ++it) {
      mbar_wait(smem_u32(&sh->a_full[s]), ph);
      tc_fence_after();
      const uint32_t w_hi = base + (uint32_t)s * w_bytes;
      const uint32_t w_lo = w_hi + (uint32_t)n_w * 128u;
      const uint32_t a_hi = tmem_a + (uint32_t)(64 * s), a_lo = a_hi + 32;
      const int k_valid = min(kGemmKChunk, P.cin - c.kc * kGemmKChunk);
      const int ksteps = (k_valid + 7) >> 3;
      if (!(P.debug & 4)) {
        if (is_main) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            if (kk < ksteps) tc_mma_tf32_ts(d_acc, a_hi + 8 * kk, make_sw128_desc(w_hi + 32u * kk), idesc, (it > 0 || kk > 0) ? 1u : 0u);
        } else {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            if (kk < ksteps) {
              tc_mma_tf32_ts(d_acc, a_lo + 8 * kk, make_sw128_desc(w_hi + 32u * kk), idesc, (it > 0 || kk > 0) ? 1u : 0u);
              tc_mma_tf32_ts(d_acc, a_hi + 8 * kk, make_sw128_desc(w_lo + 32u * kk), idesc, 1u);
            }
        }
      }
      tc_commit(smem_u32(&sh->a_empty[s]));
      c.next(kc_n);
      if (++s == kTsAStages) {
        s = 0;
        ph ^= 1u;
      }
    }
    if (n_active > 0) tc_commit(smem_u32(&sh->accum));
  } else if (warp == kTsProducers / 32 + 2 && lane == 0) {
    // ================= W loader: bulk-async copies into the stage ring =================
    StageCursor c;
    c.init(off_mask);
    int s = 0;
    uint32_t ph = 0;
    for (int it = 0; it < n_active; ++it) {
      mbar_wait(smem_u32(&sh->a_empty[s]), ph ^ 1u);
      if (!(P.debug & 2)) {
        mbar_expect_tx(smem_u32(&sh->a_full[s]), w_bytes);
        bulk_g2s(base + (uint32_t)s * w_bytes, P.w_packed + P.S.block_offset(nt, c.k, c.kc), w_bytes, smem_u32(&sh->a_full[s]));
      }
      mbar_arrive(smem_u32(&sh->a_full[s]));
      c.next(kc_n);
      if (++s == kTsAStages) {
        s = 0;
        ph ^= 1u;
      }
    }
  }
  // ================= epilogue =================
  if (warp < 4 && !(P.debug & 32)) {
    if (n_active > 0) {
      mbar_wait(smem_u32(&sh->accum), 0);
      tc_fence_after();
    }
    epilogue_phase1(P, tmem_d, acc_cols, base, n_w, nt, n_active > 0, tid);
    tc_fence_before();
  }
  if (tid < kTsProducers && !(P.debug & 32)) {
    asm volatile("bar.sync 1, %0;" ::"n"(kTsProducers) : "memory");
    epilogue_phase2(P, base, n_w, nt, row0, warp, lane, kTsProducers / 32);
  }
  __syncthreads();
  if (warp == kTsProducers / 32) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(tmem_cols) : "memory");
  }
}

// Launch helper called from fsfb_gather_gemm (gemm_tc.cu) when every column tile is <= 128 wide.
int launch_gather_gemm_ts(TcParams P, bool a_vec, cudaStream_t st) {
  const int n_w_max = P.S.n_w(0);
  const size_t w_bytes = (size_t)2 * n_w_max * 128;
  const size_t staging = (size_t)kTcRows * (((n_w_max + 31) & ~31) + 4) * 4;
  const size_t fixed = (size_t)P.koff * kTcRows * 4 + sizeof(TsShared) + 1024;
  const size_t budget = 227 * 1024;
  const int stages = kTsAStages;  // one ring: stage s = TMEM columns of A + shared-memory block of W
  P.stages = stages;
  P.data_bytes = (uint32_t)align_up(std::max((size_t)stages * w_bytes, staging), 1024);
  const size_t smem = (size_t)P.data_bytes + fixed;
  static bool attr = false;
  if (!attr) {
    FSFB_CUDA(cudaFuncSetAttribute(k_gather_gemm_ts<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget));
    FSFB_CUDA(cudaFuncSetAttribute(k_gather_gemm_ts<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget));
    attr = true;
  }
  dim3 grid((unsigned)ceil_div(P.rows, kTcRows), (unsigned)P.S.n_tiles());
  if (a_vec) {
    FSFB_LAUNCH(k_gather_gemm_ts<true>, grid, kTsThreads, smem, st, P);
  } else {
    FSFB_LAUNCH(k_gather_gemm_ts<false>, grid, kTsThreads, smem, st, P);
  }
  return FSFB_OK;
}

}  // namespace fsfb
