// gemm_tc.cu — output-stationary gather-GEMM on the 5th-gen tensor cores (tcgen05 + TMEM),
// with the Linear/BN/LN/activation/residual epilogue fused.  sm_100a only.
//
// Serves every dense contraction of the FSF forward path:
//   * SimpleSparseUNet's SubMConv3d / SparseConv3d / SparseInverseConv3d
//     (projects/configs/nuScenes/FSF_nuScenes_config.py:58-70) as
//       out[r] = epi( sum_k a[nbr[k][r]] @ w[k]^T ),  27 offsets, nbr from rulebook.cu
//   * build_mlp's Linear→norm→act blocks (projects/mmdet3d_plugin/ops/sst_ops.py:808-833)
//     as the koff == 1, nbr == identity case.
//
// CTA = one tile of 128 output rows x n_w (<= 256) output channels; 9 warps:
//   warps 0-7  A producers: gather 128 input rows x 32 floats per stage with branch-free predicated
//              128-bit loads (three stages of loads in flight per thread), split fp32 → tf32 hi/lo,
//              store into the 128B-swizzled K-major stage buffers; thread 0 also launches the
//              bulk-async copy (UBLKCP) of the pre-packed W block.  After the main loop warps 0-3
//              run the epilogue: TMEM → registers (tcgen05.ld 32x32b) → bias/LN/affine/residual/act
//              → global.
//   warp 8     lane 0 issues tcgen05.mma.kind::tf32 (M=128, N=n_w, K=8), three MMAs per K-step
//              (3xTF32: a_hi*w_hi + a_lo*w_hi + a_hi*w_lo) accumulating in TMEM, and
//              tcgen05.commit to recycle stages.
// Offsets k for which no row of the tile has a neighbour are skipped by every role.
//
// Roofline: tensor-bound for the 27-offset convolutions (2*128*n_w*32*3 flop per stage);
// HBM-bound for the plain Linear case (4*rows*(cin+cout) bytes).
#include <cstdlib>

#include "gemm_tc_ptx.cuh"

namespace fsfb {

constexpr int kTcProducers = 256;      // threads 0..255 (8 warps)
constexpr int kTcThreads = 288;        // + MMA warp

struct TcShared {  // lives after the stage buffers
  uint64_t full[kTcMaxStages];
  uint64_t empty[kTcMaxStages];
  uint64_t accum;
  uint32_t tmem_base;
  uint32_t off_mask;
};

constexpr int kRowsPerThread = kTcRows * 8 / kTcProducers;  // 16-byte chunks: 8 per row → 4 rows per thread

// DEEP = true : 1 CTA/SM, three stages of gathers in flight per thread (27-offset convolutions)
// DEEP = false: 2 CTAs/SM (<= 112 registers), two stages in flight — short K loops (Linear layers), where
//               co-resident CTAs hide each other's pipeline fill, drain and epilogue
template <bool AVEC, bool DEEP>
__global__ void __launch_bounds__(kTcThreads, DEEP ? 1 : 2) k_gather_gemm_tc(const TcParams P) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nt = blockIdx.y;
  const int n_w = P.S.n_w(nt);
  const int64_t row0 = (int64_t)blockIdx.x * kTcRows;
  const uint32_t w_bytes = (uint32_t)P.S.block_bytes(nt);
  const uint32_t stage_bytes = 2 * kStageABytes + w_bytes;
  const uint32_t base = smem_u32(smem_raw);  // dynamic shared memory starts 1024-aligned (no static smem here)
  const uint32_t s_nbr = base + P.data_bytes;  // [koff][128] i32
  TcShared* sh = reinterpret_cast<TcShared*>(smem_raw + (size_t)P.data_bytes + (size_t)P.koff * kTcRows * 4);

  // ---- setup -------------------------------------------------------------------------------
  if (tid == 0) {
    if (base & 1023u) __trap();
    for (int s = 0; s < P.stages; ++s) {
      mbar_init(smem_u32(&sh->full[s]), kTcProducers / 32);  // one elected arrival per producer warp
      mbar_init(smem_u32(&sh->empty[s]), 1);
    }
    mbar_init(smem_u32(&sh->accum), 1);
    sh->off_mask = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // two accumulators: main (a_hi*w_hi) and correction (a_lo*w_hi + a_hi*w_lo).  The tensor core
  // adds into fp32 accumulators with truncation, a bias that grows with the number of chained
  // MMAs; keeping the 2^-11-times-smaller correction terms out of the main chain cuts it 3x.
  uint32_t acc_cols = 32;
  while ((int)acc_cols < n_w) acc_cols <<= 1;
  const uint32_t tmem_cols = 2 * acc_cols;
  if (warp == kTcProducers / 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(&sh->tmem_base)),
                 "r"(tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  __syncthreads();  // barriers + off_mask initialised before the atomics below
  // neighbour tile → shared memory; which offsets have any row in this tile.
  // thread t: row t & 127, offsets (t >> 7), (t >> 7) + 2, ...  (loads of a batch are independent)
  if (tid < kTcProducers) {
    const int r_l = tid & (kTcRows - 1);
    const int64_t r = row0 + r_l < P.rows ? (P.row_order ? (int64_t)__ldg(P.row_order + row0 + r_l) : row0 + r_l) : P.rows;
    uint32_t my_mask = 0;
    constexpr int kPar = kTcProducers / kTcRows;
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
  const int kc_n = P.S.kc();
  const int n_active = __popc(off_mask) * kc_n;

  if (tid < kTcProducers && !(P.debug & 128)) {
    // ================= A producers (8 warps) =================
    const int chunk = tid & 7;   // 16-byte chunk of the 128-byte K row
    const int rbase = tid >> 3;  // rows rbase + (kTcProducers/8)*p
    auto load_stage = [&](const StageCursor& c, float4(&v)[kRowsPerThread]) {
      const int col = c.kc * kGemmKChunk + chunk * 4;
      int32_t src[kRowsPerThread];
#pragma unroll
      for (int p = 0; p < kRowsPerThread; ++p) src[p] = lds_i32(s_nbr + (uint32_t)(c.k * kTcRows + rbase + (kTcProducers / 8) * p) * 4u);
#pragma unroll
      for (int p = 0; p < kRowsPerThread; ++p) {
        const bool ok = src[p] >= 0 && col < P.cin && !(P.debug & 1);
        const float* g = P.a + (int64_t)(ok ? src[p] : 0) * P.a_stride + (ok ? col : 0);
        if (AVEC) {
          if (col + 4 <= P.cin) {  // uniform per thread
            v[p] = ldg_pred_f4(g, ok);
          } else {
            v[p].x = ldg_pred_f1(g, ok);
            v[p].y = ldg_pred_f1(g + 1, ok && col + 1 < P.cin);
            v[p].z = ldg_pred_f1(g + 2, ok && col + 2 < P.cin);
            v[p].w = ldg_pred_f1(g + 3, ok && col + 3 < P.cin);
          }
        } else {
          v[p].x = ldg_pred_f1(g, ok);
          v[p].y = ldg_pred_f1(g + 1, ok && col + 1 < P.cin);
          v[p].z = ldg_pred_f1(g + 2, ok && col + 2 < P.cin);
          v[p].w = ldg_pred_f1(g + 3, ok && col + 3 < P.cin);
        }
      }
    };
    // three stages of loads in flight: `cur` is being written, `n1`, `n2` are outstanding
    float4 cur[kRowsPerThread], n1[kRowsPerThread], n2[kRowsPerThread];
    StageCursor c_cur, c_ld;
    c_cur.init(off_mask);
    c_ld.init(off_mask);
    if (n_active > 0) load_stage(c_ld, cur);
    c_ld.next(kc_n);
    if (DEEP) {
      if (n_active > 1) load_stage(c_ld, n1);
      c_ld.next(kc_n);
    }
    int s = 0;
    uint32_t ph = 0;
    for (int it = 0; it < n_active; ++it) {
      if (DEEP) {
        if (it + 2 < n_active) load_stage(c_ld, n2);
      } else {
        if (it + 1 < n_active) load_stage(c_ld, n1);
      }
      c_ld.next(kc_n);
      if (lane == 0 && !(P.debug & 512)) mbar_wait(smem_u32(&sh->empty[s]), ph ^ 1u);  // one poller per warp
      __syncwarp();
      const uint32_t st = base + (uint32_t)s * stage_bytes;
      if (tid == 0 && !(P.debug & 2)) {
        mbar_expect_tx(smem_u32(&sh->full[s]), w_bytes);
        bulk_g2s(st + 2 * kStageABytes, P.w_packed + P.S.block_offset(nt, c_cur.k, c_cur.kc), w_bytes,
                 smem_u32(&sh->full[s]));
      }
#pragma unroll
      for (int p = 0; p < kRowsPerThread; ++p) {
        const int r = rbase + (kTcProducers / 8) * p;
        const uint32_t off = (uint32_t)r * 128u + (uint32_t)((chunk ^ (r & 7)) << 4);
        const float4 v = cur[p];
        const float4 hi = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
        const float4 lo = make_float4(tf32_lo(v.x, hi.x), tf32_lo(v.y, hi.y), tf32_lo(v.z, hi.z), tf32_lo(v.w, hi.w));
        if (!(P.debug & 8)) {
          sts_f4(st + off, hi);
          sts_f4(st + kStageABytes + off, lo);
        }
      }
      if (!(P.debug & 256)) fence_proxy_async();  // every writer orders its generic-proxy stores before the async-proxy reads
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&sh->full[s]));
#pragma unroll
      for (int p = 0; p < kRowsPerThread; ++p) {
        cur[p] = n1[p];
        if (DEEP) n1[p] = n2[p];
      }
      c_cur.next(kc_n);
      if (++s == P.stages) {
        s = 0;
        ph ^= 1u;
      }
    }
  }
  if (warp < 4 && !(P.debug & 32)) {
    // ================= epilogue (warps 0-3: TMEM lane == tile row == tid) =================
    if (n_active > 0 && !(P.debug & 128)) {
      mbar_wait(smem_u32(&sh->accum), 0);
      tc_fence_after();
    }
    epilogue_phase1(P, tmem_d, acc_cols, base, n_w, nt, n_active > 0, tid);
    tc_fence_before();
  }
  if (tid < kTcProducers && !(P.debug & 32)) {
    // phase 2 (all 8 producer warps, lanes along channels): residual + activation + coalesced stores
    asm volatile("bar.sync 1, %0;" ::"n"(kTcProducers) : "memory");
    epilogue_phase2(P, base, n_w, nt, row0, warp, lane, kTcProducers / 32);
  } else if (warp == kTcProducers / 32 && !(P.debug & 128)) {
    // ================= MMA issuer warp =================
    // warp-uniform loop, one elected lane issues: keeps descriptors in uniform registers (a UTCHMMA fed from
    // vector registers pays ~100 clk of R2UR moves, more than the math of a 128x128x8 tf32 MMA)
    const uint32_t u_tmem_d = __shfl_sync(0xffffffffu, tmem_d, 0);
    const uint32_t u_base = __shfl_sync(0xffffffffu, base, 0);
    const uint32_t u_mask = __shfl_sync(0xffffffffu, off_mask, 0);
    const uint32_t bar_full0 = __shfl_sync(0xffffffffu, smem_u32(&sh->full[0]), 0);
    const uint32_t bar_empty0 = __shfl_sync(0xffffffffu, smem_u32(&sh->empty[0]), 0);
    const uint32_t idesc = make_idesc_tf32(n_w);
    StageCursor c;
    c.init(u_mask);
    int s = 0;
    uint32_t ph = 0;
    for (int it = 0; it < n_active; ++it) {
      if (lane == 0) mbar_wait(bar_full0 + 8u * s, ph);
      __syncwarp();
      tc_fence_after();
      const uint32_t st = u_base + (uint32_t)s * stage_bytes;
      const uint32_t a_hi = st, a_lo = st + kStageABytes;
      const uint32_t w_hi = st + 2 * kStageABytes, w_lo = w_hi + (uint32_t)n_w * 128u;
      const int k_valid = min(kGemmKChunk, P.cin - c.kc * kGemmKChunk);
      const int ksteps = (k_valid + 7) >> 3;
      uint32_t elected;
      asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(elected));
      if (elected) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          if (kk < ksteps && !(P.debug & 4)) {
            const uint32_t ko = (uint32_t)kk * 32u;  // 8 tf32 = 32 bytes along K inside the swizzle row
            const uint64_t da_hi = make_sw128_desc(a_hi + ko), da_lo = make_sw128_desc(a_lo + ko);
            const uint64_t db_hi = make_sw128_desc(w_hi + ko), db_lo = make_sw128_desc(w_lo + ko);
            const uint32_t first = (it > 0 || kk > 0) ? 1u : 0u;
            tc_mma_tf32(u_tmem_d, da_hi, db_hi, idesc, first);
            tc_mma_tf32(u_tmem_d + acc_cols, da_lo, db_hi, idesc, first);
            tc_mma_tf32(u_tmem_d + acc_cols, da_hi, db_lo, idesc, 1u);
          }
        }
        tc_commit(bar_empty0 + 8u * s);
        if (it == n_active - 1) tc_commit(smem_u32(&sh->accum));
      }
      __syncwarp();
      c.next(kc_n);
      if (++s == P.stages) {
        s = 0;
        ph ^= 1u;
      }
    }
  }
  __syncthreads();
  if (warp == kTcProducers / 32) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(tmem_cols) : "memory");
  }
}

}  // namespace fsfb

namespace fsfb {
static int gather_gemm_impl(const float* a, int64_t a_rows, int cin, int64_t a_stride, const int32_t* nbr,
                            const int32_t* row_order, int koff, int64_t rows, const void* w_packed, int cout, const float* bias,
                            int norm, const float* norm_w, const float* norm_b, float eps,
                            const float* residual, int64_t residual_stride, int act, float* out,
                            int64_t out_stride, int splits, void* workspace, size_t workspace_bytes, const float* host_bias,
                            const float* host_norm_w, const float* host_norm_b, void* stream, bool a_split = false) {
  const bool nbr_ro = (act & FSFB_NBR_ROW_ORDERED) != 0;   // include/fsf_b200.h: the table is permuted into the row order
  act &= ~FSFB_NBR_ROW_ORDERED;
  FSFB_CHECK_ARG(!nbr_ro || (a_split && nbr && row_order), "gather_gemm: FSFB_NBR_ROW_ORDERED needs fsfb_gather_gemm_split with a row order");
  int rc = check_epilogue(cout, bias, norm, norm_w, norm_b, act, "gather_gemm");
  if (rc != FSFB_OK) return rc;
  FSFB_CHECK_ARG(rows >= 0 && a_rows >= 0 && a_rows < (1ll << 31) && cin >= 1 && a_stride >= cin &&
                     out_stride >= cout,
                 "gather_gemm: bad shape rows=%lld a_rows=%lld cin=%d cout=%d", (long long)rows,
                 (long long)a_rows, cin, cout);
  FSFB_CHECK_ARG(koff >= 1 && koff <= kTcMaxOff, "gather_gemm: koff=%d unsupported (1..%d)", koff, kTcMaxOff);
  FSFB_CHECK_ARG(nbr || koff == 1, "gather_gemm: koff > 1 needs a neighbour table");
  FSFB_CHECK_ARG(norm != FSFB_NORM_LAYERNORM || cout <= kGemmNTile || (koff == 1 && !nbr && splits > 1 && cout <= 1024),
                 "gather_gemm: fused LayerNorm needs cout <= %d (use fsfb_rownorm_act after the GEMM)",
                 kGemmNTile);
  if (rows == 0) return FSFB_OK;
  FSFB_CHECK_ARG(w_packed && out && (a || a_rows == 0), "gather_gemm: null pointer");
  FSFB_CHECK_ARG(((uintptr_t)w_packed & 15) == 0, "gather_gemm: packed weights must be 16-byte aligned");
  TcParams P;
  P.a = a;
  P.a_rows = a_rows;
  P.cin = cin;
  P.a_stride = a_stride;
  P.nbr = nbr;
  P.row_order = row_order;
  P.nbr_ro = nbr_ro ? 1 : 0;
  P.koff = koff;
  P.rows = rows;
  P.w_packed = (const unsigned char*)w_packed;
  P.S = GemmShape{koff, cin, cout};
  P.E = Epilogue{bias, norm, norm_w, norm_b, eps, residual, residual_stride, act};
  P.out = out;
  P.out_stride = out_stride;
  const bool a_vec = ((uintptr_t)a % 16 == 0) && (a_stride % 4 == 0);
  P.out_vec = ((uintptr_t)out % 16 == 0) && (out_stride % 4 == 0);
  {
    const char* dbg = getenv("FSFB_GEMM_DEBUG");
    P.debug = dbg ? atoi(dbg) : 0;
  }
  const int n_w_max = P.S.n_w(0);
  {
    // persistent A-through-TMEM kernel (gemm_ts.cu) whenever the output splits into column tiles of <= 128
    // channels; the kernel below remains for widths like 131 / 144 (one 256-wide tile) and fused LayerNorms wider
    // than 128.  FSFB_GEMM_TS=0 forces the kernel below (A/B experiments).
    static const int ss_mode = [] { const char* e = getenv("FSFB_GEMM_SS"); return e ? atoi(e) : 1; }();
    if (ss_mode && !a_split && (splits <= 1 || (koff == 1 && !nbr))) {  // dense Linear: one CTA per row tile (gemm_lin.cu)
      // (a Linear layer has one "offset": splits > 1 there means K-chunk ranges, which only this kernel implements)
      const int rc_lin = launch_linear_ss(P, a_vec, splits, (float*)workspace, workspace_bytes, (cudaStream_t)stream);
      if (rc_lin != 1) return rc_lin;
      FSFB_CHECK_ARG(splits <= 1, "gather_gemm: K split of a Linear layer needs the row-tile kernel (rows >= 1024, FSFB_GEMM_LIN != 0)");
    }
    if (ss_mode) {  // fp16-split operands from shared memory (gemm_ss.cu); 1 = shape not served there
      const int rc_ss = launch_gather_gemm_ss(P, a_vec, a_split, (float*)workspace, workspace_bytes, splits, host_bias, host_norm_w,
                                              host_norm_b, (cudaStream_t)stream);
      if (rc_ss != 1) return rc_ss;
    }
    if (a_split) {
      set_error("gather_gemm_split: shape koff=%d cin=%d cout=%d is not served by the fp16-split kernel (or FSFB_GEMM_SS/F16=0)", koff, cin, cout);
      return FSFB_ERR_BADARG;
    }
    static const int ts_mode = [] { const char* e = getenv("FSFB_GEMM_TS"); return e ? atoi(e) : 1; }();
    const int n_pad = P.S.n_pad();
    // (31 or 32 offsets: two neighbour tables no longer fit next to the rings in shared memory)
    const bool ts_ok = koff <= 30 && (n_pad <= 128 || (n_pad % 128 == 0 && norm != FSFB_NORM_LAYERNORM));
    if (ts_mode && ts_ok)
      return launch_gather_gemm_ts(P, a_vec, (float*)workspace, workspace_bytes, splits, host_bias, host_norm_w, host_norm_b,
                                   (cudaStream_t)stream);
    FSFB_CHECK_ARG(splits <= 1, "gather_gemm: offset splits need cout <= 128 or a multiple of 128 (cout=%d)", cout);
  }
  const size_t stage_bytes = 2 * (size_t)kStageABytes + (size_t)2 * n_w_max * 128;
  const size_t fixed = (size_t)koff * kTcRows * 4 + sizeof(TcShared) + 1024 /* alignment slack */;
  const size_t budget = 227 * 1024;
  int stages = (int)std::min<size_t>(kTcMaxStages, (budget - fixed) / stage_bytes);
  FSFB_CHECK_ARG(stages >= 1, "gather_gemm: tile does not fit shared memory");
  const int64_t total_iters = (int64_t)koff * P.S.kc();
  if (total_iters < stages) stages = (int)std::max<int64_t>(1, total_iters);
  // short K loops: shrink the ring so two CTAs share an SM (TMEM: 2 x 256 columns, smem: 2 x <= 113 KB)
  const bool deep = total_iters > 8 || n_w_max > 128;
  if (!deep) {
    const size_t half = (budget - 2 * 1024) / 2;
    while (stages > 1 && (size_t)stages * stage_bytes + fixed > half) --stages;
  }
  P.stages = stages;
  const size_t staging = (size_t)kTcRows * (((n_w_max + 31) & ~31) + 4) * 4;  // epilogue row staging reuses the stage buffers
  P.data_bytes = (uint32_t)align_up(std::max((size_t)stages * stage_bytes, staging), 1024);
  const size_t smem = (size_t)P.data_bytes + fixed;
  static bool attr = false;
  if (!attr) {
    FSFB_CUDA(cudaFuncSetAttribute(k_gather_gemm_tc<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget));
    FSFB_CUDA(cudaFuncSetAttribute(k_gather_gemm_tc<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget));
    FSFB_CUDA(cudaFuncSetAttribute(k_gather_gemm_tc<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget));
    FSFB_CUDA(cudaFuncSetAttribute(k_gather_gemm_tc<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget));
    attr = true;
  }
  dim3 grid((unsigned)ceil_div(rows, kTcRows), (unsigned)P.S.n_tiles());
  cudaStream_t st = (cudaStream_t)stream;
  if (deep) {
    if (a_vec) FSFB_LAUNCH((k_gather_gemm_tc<true, true>), grid, kTcThreads, smem, st, P);
    else FSFB_LAUNCH((k_gather_gemm_tc<false, true>), grid, kTcThreads, smem, st, P);
  } else {
    if (a_vec) FSFB_LAUNCH((k_gather_gemm_tc<true, false>), grid, kTcThreads, smem, st, P);
    else FSFB_LAUNCH((k_gather_gemm_tc<false, false>), grid, kTcThreads, smem, st, P);
  }
  return FSFB_OK;
}
}  // namespace fsfb

extern "C" int fsfb_gather_gemm(const float* a, int64_t a_rows, int cin, int64_t a_stride, const int32_t* nbr,
                                const int32_t* row_order, int koff, int64_t rows, const void* w_packed, int cout, const float* bias,
                                int norm, const float* norm_w, const float* norm_b, float eps,
                                const float* residual, int64_t residual_stride, int act, float* out,
                                int64_t out_stride, void* stream) {
  return fsfb::gather_gemm_impl(a, a_rows, cin, a_stride, nbr, row_order, koff, rows, w_packed, cout, bias, norm, norm_w, norm_b,
                                eps, residual, residual_stride, act, out, out_stride, 1, nullptr, 0, nullptr, nullptr, nullptr, stream);
}

extern "C" int fsfb_gather_gemm_splitk_bytes(int64_t rows, int cout, int splits, size_t* bytes) {
  using namespace fsfb;
  FSFB_CHECK_ARG(bytes && rows >= 0 && cout >= 1 && splits >= 1, "gather_gemm_splitk_bytes: bad argument");
  const size_t cpad = (size_t)((cout + 127) / 128 * 128);
  *bytes = splits > 1 ? (size_t)splits * (size_t)rows * cpad * sizeof(float) : 0;
  return FSFB_OK;
}

extern "C" int fsfb_gather_gemm_splitk(const float* a, int64_t a_rows, int cin, int64_t a_stride, const int32_t* nbr,
                                       const int32_t* row_order, int koff, int64_t rows, const void* w_packed, int cout,
                                       const float* bias, int norm, const float* norm_w, const float* norm_b, float eps,
                                       const float* residual, int64_t residual_stride, int act, float* out,
                                       int64_t out_stride, int splits, void* workspace, size_t workspace_bytes, void* stream) {
  using namespace fsfb;
  FSFB_CHECK_ARG(splits >= 1 && (splits <= koff || (koff == 1 && !nbr && splits <= (cin + 31) / 32)),
                 "gather_gemm_splitk: splits=%d must be in 1..koff (Linear: 1..K chunks)", splits);
  FSFB_CHECK_ARG(splits == 1 || ((uintptr_t)workspace & 15) == 0, "gather_gemm_splitk: workspace must be 16-byte aligned");
  return gather_gemm_impl(a, a_rows, cin, a_stride, nbr, row_order, koff, rows, w_packed, cout, bias, norm, norm_w, norm_b, eps,
                          residual, residual_stride, act, out, out_stride, splits, workspace, workspace_bytes, nullptr, nullptr, nullptr, stream);
}

extern "C" int fsfb_gather_gemm_hv(const float* a, int64_t a_rows, int cin, int64_t a_stride, const int32_t* nbr,
                                   const int32_t* row_order, int koff, int64_t rows, const void* w_packed, int cout,
                                   const float* bias, int norm, const float* norm_w, const float* norm_b, float eps,
                                   const float* residual, int64_t residual_stride, int act, float* out, int64_t out_stride,
                                   int splits, void* workspace, size_t workspace_bytes, const float* host_bias,
                                   const float* host_norm_w, const float* host_norm_b, void* stream) {
  using namespace fsfb;
  FSFB_CHECK_ARG(splits >= 1 && splits <= koff, "gather_gemm_hv: splits=%d must be in 1..koff", splits);
  FSFB_CHECK_ARG(splits == 1 || ((uintptr_t)workspace & 15) == 0, "gather_gemm_hv: workspace must be 16-byte aligned");
  return gather_gemm_impl(a, a_rows, cin, a_stride, nbr, row_order, koff, rows, w_packed, cout, bias, norm, norm_w, norm_b, eps,
                          residual, residual_stride, act, out, out_stride, splits, workspace, workspace_bytes, host_bias,
                          host_norm_w, host_norm_b, stream);
}

extern "C" int fsfb_gather_gemm_split(const void* a_split, int64_t a_rows, int cin, const int32_t* nbr, const int32_t* row_order, int koff,
                                      int64_t rows, const void* w_packed, int cout, const float* bias, int norm, const float* norm_w,
                                      const float* norm_b, float eps, const float* residual, int64_t residual_stride, int act, float* out,
                                      int64_t out_stride, int splits, void* workspace, size_t workspace_bytes, const float* host_bias,
                                      const float* host_norm_w, const float* host_norm_b, void* stream) {
  using namespace fsfb;
  FSFB_CHECK_ARG(splits >= 1 && splits <= koff, "gather_gemm_split: splits=%d must be in 1..koff", splits);
  FSFB_CHECK_ARG(splits == 1 || ((uintptr_t)workspace & 15) == 0, "gather_gemm_split: workspace must be 16-byte aligned");
  FSFB_CHECK_ARG(cin % 32 == 0, "gather_gemm_split: cin=%d must be a multiple of 32", cin);
  return gather_gemm_impl(reinterpret_cast<const float*>(a_split), a_rows, cin, cin, nbr, row_order, koff, rows, w_packed, cout, bias, norm,
                          norm_w, norm_b, eps, residual, residual_stride, act, out, out_stride, splits, workspace, workspace_bytes,
                          host_bias, host_norm_w, host_norm_b, stream, true);
}
