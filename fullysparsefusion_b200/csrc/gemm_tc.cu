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
// CTA = one tile of 128 output rows x n_w (<= 256) output channels; 5 warps:
//   warps 0-3  A producers: gather 128 input rows x 32 floats per stage with 128-bit loads,
//              split fp32 → tf32 hi/lo, store into the 128B-swizzled K-major stage buffers;
//              thread 0 also launches the bulk-async copy (UBLKCP) of the pre-packed W block.
//              After the main loop the same warps run the epilogue: TMEM → registers
//              (tcgen05.ld 32x32b) → bias/LN/affine/residual/act → global.
//   warp 4     lane 0 issues tcgen05.mma.kind::tf32 (M=128, N=n_w, K=8), three MMAs per K-step
//              (3xTF32: a_hi*w_hi + a_lo*w_hi + a_hi*w_lo) accumulating in TMEM, and
//              tcgen05.commit to recycle stages.
// Offsets k for which no row of the tile has a neighbour are skipped by every role.
//
// Roofline: tensor-bound for the 27-offset convolutions (2*128*n_w*32*3 flop per stage);
// HBM-bound for the plain Linear case (4*rows*(cin+cout) bytes).
#include "gemm_common.cuh"

namespace fsfb {

constexpr int kTcRows = 128;           // UMMA M
constexpr int kTcProducers = 128;      // threads 0..127
constexpr int kTcThreads = 160;        // + MMA warp
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

struct TcParams {
  const float* a;
  int64_t a_rows;
  int cin;
  int64_t a_stride;
  const int32_t* nbr;
  int koff;
  int64_t rows;
  const unsigned char* w_packed;
  GemmShape S;
  Epilogue E;
  float* out;
  int64_t out_stride;
  int stages;
  int a_vec;    // 1: rows of `a` are 16-byte aligned → 128-bit loads
  int out_vec;  // 1: rows of `out` are 16-byte aligned → 128-bit stores
};

struct TcShared {  // lives after the stage buffers
  uint64_t full[kTcMaxStages];
  uint64_t empty[kTcMaxStages];
  uint64_t accum;
  uint32_t tmem_base;
  uint32_t off_mask;
};

__global__ void __launch_bounds__(kTcThreads, 1) k_gather_gemm_tc(const TcParams P) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nt = blockIdx.y;
  const int n_w = P.S.n_w(nt);
  const int64_t row0 = (int64_t)blockIdx.x * kTcRows;
  const uint32_t w_bytes = (uint32_t)P.S.block_bytes(nt);
  const uint32_t stage_bytes = 2 * kStageABytes + w_bytes;
  // 1024-byte aligned carve-up
  unsigned char* base = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  int32_t* s_nbr = reinterpret_cast<int32_t*>(base + (size_t)P.stages * stage_bytes);  // [koff][128]
  TcShared* sh = reinterpret_cast<TcShared*>(s_nbr + P.koff * kTcRows);

  // ---- setup -------------------------------------------------------------------------------
  if (tid == 0) {
    for (int s = 0; s < P.stages; ++s) {
      mbar_init(smem_u32(&sh->full[s]), kTcProducers);
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
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(&sh->tmem_base)),
                 "r"(tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  __syncthreads();  // barriers + off_mask initialised before the atomics below
  // neighbour tile → shared memory; which offsets have any row in this tile
  uint32_t my_mask = 0;
  if (tid < kTcProducers) {
    const int64_t r = row0 + tid;
    for (int k = 0; k < P.koff; ++k) {
      int32_t src = -1;
      if (r < P.rows) {
        src = P.nbr ? __ldg(P.nbr + (int64_t)k * P.rows + r) : (int32_t)r;
        if (src >= P.a_rows) src = -1;
      }
      s_nbr[k * kTcRows + tid] = src;
      my_mask |= (src >= 0 ? 1u : 0u) << k;
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

  if (warp < 4) {
    // ================= A producers =================
    const int chunk = tid & 7;       // 16-byte chunk of the 128-byte K row
    const int rbase = tid >> 3;      // rows rbase + 16*p
    float4 cur[8], nxt[8];
    auto load_stage = [&](int k, int kchunk, float4(&v)[8]) {
      const int col = kchunk * kGemmKChunk + chunk * 4;
#pragma unroll
      for (int p = 0; p < 8; ++p) {
        const int32_t src = s_nbr[k * kTcRows + rbase + 16 * p];
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        if (src >= 0 && col < P.cin) {
          const float* g = P.a + (int64_t)src * P.a_stride + col;
          if (P.a_vec && col + 4 <= P.cin) {
            t = __ldg(reinterpret_cast<const float4*>(g));
          } else {
            t.x = __ldg(g);
            if (col + 1 < P.cin) t.y = __ldg(g + 1);
            if (col + 2 < P.cin) t.z = __ldg(g + 2);
            if (col + 3 < P.cin) t.w = __ldg(g + 3);
          }
        }
        v[p] = t;
      }
    };
    // iteration cursor over (active offset k, k-chunk)
    uint32_t rem = off_mask;
    int k_cur = rem ? __ffs(rem) - 1 : 0, kc_cur = 0;
    auto advance = [&](uint32_t& m, int& k, int& kchunk) {
      if (++kchunk == kc_n) {
        kchunk = 0;
        m &= m - 1;
        k = m ? __ffs(m) - 1 : 0;
      }
    };
    if (n_active > 0) load_stage(k_cur, kc_cur, cur);
    for (int it = 0; it < n_active; ++it) {
      const int s = it % P.stages;
      const uint32_t ph = (uint32_t)(it / P.stages) & 1u;
      // prefetch the next stage's rows into registers before blocking on the ring
      uint32_t rem_n = rem;
      int k_n = k_cur, kc_nx = kc_cur;
      advance(rem_n, k_n, kc_nx);
      if (it + 1 < n_active) load_stage(k_n, kc_nx, nxt);
      mbar_wait(smem_u32(&sh->empty[s]), ph ^ 1u);
      unsigned char* st = base + (size_t)s * stage_bytes;
      if (tid == 0) {
        mbar_expect_tx(smem_u32(&sh->full[s]), w_bytes);
        bulk_g2s(smem_u32(st + 2 * kStageABytes), P.w_packed + P.S.block_offset(nt, k_cur, kc_cur), w_bytes,
                 smem_u32(&sh->full[s]));
      }
#pragma unroll
      for (int p = 0; p < 8; ++p) {
        const int r = rbase + 16 * p;
        const uint32_t off = (uint32_t)r * 128u + (uint32_t)((chunk ^ (r & 7)) << 4);
        const float4 v = cur[p];
        const float4 hi = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
        const float4 lo = make_float4(tf32_lo(v.x, hi.x), tf32_lo(v.y, hi.y), tf32_lo(v.z, hi.z), tf32_lo(v.w, hi.w));
        *reinterpret_cast<float4*>(st + off) = hi;
        *reinterpret_cast<float4*>(st + kStageABytes + off) = lo;
      }
      fence_proxy_async();
      mbar_arrive(smem_u32(&sh->full[s]));
#pragma unroll
      for (int p = 0; p < 8; ++p) cur[p] = nxt[p];
      rem = rem_n;
      k_cur = k_n;
      kc_cur = kc_nx;
    }

    // ================= epilogue =================
    if (n_active > 0) {
      mbar_wait(smem_u32(&sh->accum), 0);
      tc_fence_after();
    }
    const int64_t r = row0 + tid;  // TMEM lane == tile row == tid
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
    for (int cb = 0; cb < c_n; cb += 32) {
      if (n_active > 0) {
        tc_ld32(t_row + cb, v);  // warp-collective: executed by all lanes, valid row or not
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0.f;
      }
      if (r < P.rows) {
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
            if (E.residual) x += __ldg(E.residual + r * E.residual_stride + c);
            v[j] = apply_act(x, E.act);
          }
        }
        float* o = P.out + r * P.out_stride + c0 + cb;
        if (P.out_vec && cb + 32 <= c_n) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (cb + j < c_n) o[j] = v[j];
        }
      }
    }
    tc_fence_before();
  } else if (lane == 0) {
    // ================= MMA issuer =================
    const uint32_t idesc = make_idesc_tf32(n_w);
    uint32_t rem = off_mask;
    int k_cur = rem ? __ffs(rem) - 1 : 0, kc_cur = 0;
    for (int it = 0; it < n_active; ++it) {
      const int s = it % P.stages;
      const uint32_t ph = (uint32_t)(it / P.stages) & 1u;
      mbar_wait(smem_u32(&sh->full[s]), ph);
      tc_fence_after();
      const uint32_t st = smem_u32(base + (size_t)s * stage_bytes);
      const uint32_t a_hi = st, a_lo = st + kStageABytes;
      const uint32_t w_hi = st + 2 * kStageABytes, w_lo = w_hi + (uint32_t)n_w * 128u;
      const int k_valid = min(kGemmKChunk, P.cin - kc_cur * kGemmKChunk);
      const int ksteps = (k_valid + 7) >> 3;
      for (int kk = 0; kk < ksteps; ++kk) {
        const uint32_t ko = (uint32_t)kk * 32u;  // 8 tf32 = 32 bytes along K inside the swizzle row
        const uint64_t da_hi = make_sw128_desc(a_hi + ko), da_lo = make_sw128_desc(a_lo + ko);
        const uint64_t db_hi = make_sw128_desc(w_hi + ko), db_lo = make_sw128_desc(w_lo + ko);
        const uint32_t first = (it > 0 || kk > 0) ? 1u : 0u;
        tc_mma_tf32(tmem_d, da_hi, db_hi, idesc, first);
        tc_mma_tf32(tmem_d + acc_cols, da_lo, db_hi, idesc, first);
        tc_mma_tf32(tmem_d + acc_cols, da_hi, db_lo, idesc, 1u);
      }
      tc_commit(smem_u32(&sh->empty[s]));
      if (++kc_cur == kc_n) {
        kc_cur = 0;
        rem &= rem - 1;
        k_cur = rem ? __ffs(rem) - 1 : 0;
      }
    }
    if (n_active > 0) tc_commit(smem_u32(&sh->accum));
    (void)k_cur;
  }
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(tmem_cols) : "memory");
  }
}

}  // namespace fsfb

extern "C" int fsfb_gather_gemm(const float* a, int64_t a_rows, int cin, int64_t a_stride, const int32_t* nbr,
                                int koff, int64_t rows, const void* w_packed, int cout, const float* bias,
                                int norm, const float* norm_w, const float* norm_b, float eps,
                                const float* residual, int64_t residual_stride, int act, float* out,
                                int64_t out_stride, void* stream) {
  using namespace fsfb;
  int rc = check_epilogue(cout, bias, norm, norm_w, norm_b, act, "gather_gemm");
  if (rc != FSFB_OK) return rc;
  FSFB_CHECK_ARG(rows >= 0 && a_rows >= 0 && a_rows < (1ll << 31) && cin >= 1 && a_stride >= cin &&
                     out_stride >= cout,
                 "gather_gemm: bad shape rows=%lld a_rows=%lld cin=%d cout=%d", (long long)rows,
                 (long long)a_rows, cin, cout);
  FSFB_CHECK_ARG(koff >= 1 && koff <= kTcMaxOff, "gather_gemm: koff=%d unsupported (1..%d)", koff, kTcMaxOff);
  FSFB_CHECK_ARG(nbr || koff == 1, "gather_gemm: koff > 1 needs a neighbour table");
  FSFB_CHECK_ARG(norm != FSFB_NORM_LAYERNORM || cout <= kGemmNTile,
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
  P.koff = koff;
  P.rows = rows;
  P.w_packed = (const unsigned char*)w_packed;
  P.S = GemmShape{koff, cin, cout};
  P.E = Epilogue{bias, norm, norm_w, norm_b, eps, residual, residual_stride, act};
  P.out = out;
  P.out_stride = out_stride;
  P.a_vec = ((uintptr_t)a % 16 == 0) && (a_stride % 4 == 0);
  P.out_vec = ((uintptr_t)out % 16 == 0) && (out_stride % 4 == 0);
  const int n_w_max = P.S.n_w(0);
  const size_t stage_bytes = 2 * (size_t)kStageABytes + (size_t)2 * n_w_max * 128;
  const size_t fixed = (size_t)koff * kTcRows * 4 + sizeof(TcShared) + 1024 /* alignment slack */;
  const size_t budget = 227 * 1024;
  int stages = (int)std::min<size_t>(kTcMaxStages, (budget - fixed) / stage_bytes);
  FSFB_CHECK_ARG(stages >= 1, "gather_gemm: tile does not fit shared memory");
  const int64_t total_iters = (int64_t)koff * P.S.kc();
  if (total_iters < stages) stages = (int)std::max<int64_t>(1, total_iters);
  P.stages = stages;
  const size_t smem = (size_t)stages * stage_bytes + fixed;
  static bool attr = false;
  if (!attr) {
    FSFB_CUDA(cudaFuncSetAttribute(k_gather_gemm_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget));
    attr = true;
  }
  dim3 grid((unsigned)ceil_div(rows, kTcRows), (unsigned)P.S.n_tiles());
  FSFB_LAUNCH(k_gather_gemm_tc, grid, kTcThreads, smem, (cudaStream_t)stream, P);
  return FSFB_OK;
}
