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

__device__ __forceinline__ void tc_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
      "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
      "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]),
      "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]),
      "r"(v[31])
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
      mbar_init(smem_u32(&sh->a_full[s]), 4 + 1);  // the 4 warps of the stage's producer group + the W loader (with its tx bytes)
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
    const int64_t r = row0 + r_l < P.rows ? (P.row_order ? (int64_t)__ldg(P.row_order + row0 + r_l) : row0 + r_l) : P.rows;
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
    // ================= A producers: stage-striped warp groups =================
    // group g = warps 4g..4g+3 (one per TMEM lane quarter) produces stages it = g, g+4, g+8, ... into ring slot g:
    // a thread owns one ROW of the tile and gathers its whole 128-byte K chunk (8 predicated 128-bit loads), so the
    // per-stage bookkeeping (barrier wait/arrive, cursor, index lookup, addressing) is paid by 4 warps instead of
    // 16 and every group has four stage-times to hide its gather latency.
    const int grp = warp >> 2, q = warp & 3;
    const int row = 32 * q + lane;
    const uint32_t ta = tmem_a + ((uint32_t)(32 * q) << 16) + (uint32_t)(64 * grp);
    const uint32_t full_bar = smem_u32(&sh->a_full[grp]), empty_bar = smem_u32(&sh->a_empty[grp]);
    auto load_stage = [&](const StageCursor& c, float4(&v)[8]) {
      const int32_t src = lds_i32(s_nbr + (uint32_t)(c.k * kTcRows + row) * 4u);
      const int col0 = c.kc * kGemmKChunk;
      const float* g0 = P.a + (int64_t)(src >= 0 ? src : 0) * P.a_stride + col0;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int col = col0 + 4 * j;
        const bool ok = src >= 0 && col < P.cin && !(P.debug & 1);
        const float* g = ok ? g0 + 4 * j : P.a;
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
    float4 cur[8];
    StageCursor c_ld;
    c_ld.init(off_mask);
    for (int j = 0; j < grp; ++j) c_ld.next(kc_n);
    if (grp < n_active) load_stage(c_ld, cur);
    uint32_t ph = 0;
    for (int it = grp; it < n_active; it += kTsAStages) {
      if (lane == 0) mbar_wait(empty_bar, ph ^ 1u);
      __syncwarp();
      tc_fence_after();
      // hi = x with the 13 low mantissa bits cleared (what the tensor core keeps of a tf32 operand),
      // lo = x - hi exactly (the MMA truncates it to tf32 itself): 2 ALU ops per element
      uint32_t t32[32];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        t32[4 * j + 0] = __float_as_uint(cur[j].x) & 0xFFFFE000u;
        t32[4 * j + 1] = __float_as_uint(cur[j].y) & 0xFFFFE000u;
        t32[4 * j + 2] = __float_as_uint(cur[j].z) & 0xFFFFE000u;
        t32[4 * j + 3] = __float_as_uint(cur[j].w) & 0xFFFFE000u;
      }
      if (!(P.debug & 8)) tc_st32(ta, t32);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        t32[4 * j + 0] = __float_as_uint(cur[j].x - __uint_as_float(t32[4 * j + 0]));
        t32[4 * j + 1] = __float_as_uint(cur[j].y - __uint_as_float(t32[4 * j + 1]));
        t32[4 * j + 2] = __float_as_uint(cur[j].z - __uint_as_float(t32[4 * j + 2]));
        t32[4 * j + 3] = __float_as_uint(cur[j].w - __uint_as_float(t32[4 * j + 3]));
      }
      if (!(P.debug & 8)) {
        tc_st32(ta + 32, t32);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(full_bar);
      // this group's next stage is four stage-times away: its gathers fly while the other groups work
#pragma unroll
      for (int j = 0; j < kTsAStages; ++j) c_ld.next(kc_n);
      if (it + kTsAStages < n_active) load_stage(c_ld, cur);
      ph ^= 1u;
    }
  } else if (warp == kTsProducers / 32 || warp == kTsProducers / 32 + 1) {
    // ================= two MMA issuer warps =================
    // warp 16: main accumulator (a_hi * w_hi); warp 17: correction accumulator (a_lo * w_hi + a_hi * w_lo).
    // The whole warp runs the loop with warp-uniform operands (values laundered through __shfl_sync so ptxas keeps
    // descriptors / TMEM addresses in uniform registers) and one elected lane issues: a UTCHMMA whose operands
    // come from vector registers costs ~100 clk of R2UR moves per instruction, more than the 69 clk of math of a
    // 128x128x8 tf32 MMA (measured: run time was independent of N).
    const bool is_main = warp == kTsProducers / 32;
    const uint32_t u_tmem_d = __shfl_sync(0xffffffffu, tmem_d, 0);
    const uint32_t u_tmem_a = __shfl_sync(0xffffffffu, tmem_a, 0);
    const uint32_t u_base = __shfl_sync(0xffffffffu, base, 0);
    const uint32_t u_mask = __shfl_sync(0xffffffffu, off_mask, 0);
    const uint32_t idesc = make_idesc_tf32((P.debug & 1024) ? 16 : n_w);
    const uint32_t d_acc = is_main ? u_tmem_d : u_tmem_d + acc_cols;
    const uint32_t bar_full0 = __shfl_sync(0xffffffffu, smem_u32(&sh->a_full[0]), 0);
    const uint32_t bar_empty0 = __shfl_sync(0xffffffffu, smem_u32(&sh->a_empty[0]), 0);
    StageCursor c;
    c.init(u_mask);
    int s = 0;
    uint32_t ph = 0;
    for (int it = 0; it < n_active; ++it) {
      if (lane == 0) mbar_wait(bar_full0 + 8u * s, ph);
      __syncwarp();
      tc_fence_after();
      const uint32_t w_hi = u_base + (uint32_t)s * w_bytes;
      const uint32_t w_lo = w_hi + (uint32_t)n_w * 128u;
      const uint32_t a_hi = u_tmem_a + (uint32_t)(64 * s), a_lo = a_hi + 32;
      const int k_valid = min(kGemmKChunk, P.cin - c.kc * kGemmKChunk);
      const int ksteps = (k_valid + 7) >> 3;
      uint32_t elected;
      asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(elected));
      if (elected) {
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
        tc_commit(bar_empty0 + 8u * s);
        if (it == n_active - 1) tc_commit(smem_u32(&sh->accum));
      }
      __syncwarp();
      c.next(kc_n);
      if (++s == kTsAStages) {
        s = 0;
        ph ^= 1u;
      }
    }
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
