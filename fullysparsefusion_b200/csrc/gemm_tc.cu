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

#include "gemm_common.cuh"

namespace fsfb {

constexpr int kTcRows = 128;           // UMMA M
constexpr int kTcProducers = 256;      // threads 0..255 (8 warps)
constexpr int kTcThreads = 288;        // + MMA warp
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
  int out_vec;  // 1: rows of `out` are 16-byte aligned → 128-bit stores
  uint32_t data_bytes;  // stage ring (or epilogue staging, whichever is larger); nbr tile + barriers follow
  int debug;    // FSFB_GEMM_DEBUG bits (profiling experiments only): 1 no A loads, 2 no W copy, 4 no MMA, 8 no A stores
};

struct TcShared {  // lives after the stage buffers
  uint64_t full[kTcMaxStages];
  uint64_t empty[kTcMaxStages];
  uint64_t accum;
  uint32_t tmem_base;
  uint32_t off_mask;
};

// shared-memory accessors on 32-bit shared addresses (keeps the accesses in the shared state space:
// a pointer recovered from integer arithmetic would compile to generic LD/ST)
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

constexpr int kRowsPerThread = kTcRows * 8 / kTcProducers;  // 16-byte chunks: 8 per row → 4 rows per thread

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
    const int64_t r = row0 + r_l;
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
    tc_fence_before();
  }
  if (tid < kTcProducers && !(P.debug & 32)) {
    // phase 2 (all 8 producer warps, lanes along channels): residual + activation + coalesced stores
    asm volatile("bar.sync 1, %0;" ::"n"(kTcProducers) : "memory");
    const Epilogue& E = P.E;
    const int c0 = nt * kGemmNTile;
    const int c_n = min(n_w, P.S.cout - c0);
    const bool res_vec = E.residual && ((uintptr_t)E.residual % 16 == 0) && (E.residual_stride % 4 == 0);
    for (int rl = warp; rl < kTcRows; rl += kTcProducers / 32) {
      const int64_t r = row0 + rl;
      if (r >= P.rows || (P.debug & 16)) break;
      const uint32_t srow = base + (uint32_t)rl * (uint32_t)(((n_w + 31) & ~31) + 4) * 4u;
      float* o = P.out + r * P.out_stride + c0;
      const float* res = E.residual ? E.residual + r * E.residual_stride + c0 : nullptr;
      if (P.out_vec && (c_n & 3) == 0 && (!E.residual || res_vec)) {
        for (int c = lane * 4; c < c_n; c += 128) {
          float4 x;
          asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w) : "r"(srow + c * 4));
          if (res) {
            const float4 q = __ldg(reinterpret_cast<const float4*>(res + c));
            x.x += q.x; x.y += q.y; x.z += q.z; x.w += q.w;
          }
          x.x = apply_act(x.x, E.act); x.y = apply_act(x.y, E.act); x.z = apply_act(x.z, E.act); x.w = apply_act(x.w, E.act);
          *reinterpret_cast<float4*>(o + c) = x;
        }
      } else {
        for (int c = lane; c < c_n; c += 32) {
          float x;
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"(srow + c * 4));
          if (res) x += __ldg(res + c);
          o[c] = apply_act(x, E.act);
        }
      }
    }
  } else if (warp == kTcProducers / 32 && lane == 0 && !(P.debug & 128)) {
    // ================= MMA issuer =================
    const uint32_t idesc = make_idesc_tf32(n_w);
    StageCursor c;
    c.init(off_mask);
    int s = 0;
    uint32_t ph = 0;
    for (int it = 0; it < n_active; ++it) {
      mbar_wait(smem_u32(&sh->full[s]), ph);
      tc_fence_after();
      const uint32_t st = base + (uint32_t)s * stage_bytes;
      const uint32_t a_hi = st, a_lo = st + kStageABytes;
      const uint32_t w_hi = st + 2 * kStageABytes, w_lo = w_hi + (uint32_t)n_w * 128u;
      const int k_valid = min(kGemmKChunk, P.cin - c.kc * kGemmKChunk);
      const int ksteps = (k_valid + 7) >> 3;
      for (int kk = 0; kk < ksteps && !(P.debug & 4); ++kk) {
        const uint32_t ko = (uint32_t)kk * 32u;  // 8 tf32 = 32 bytes along K inside the swizzle row
        const uint64_t da_hi = make_sw128_desc(a_hi + ko), da_lo = make_sw128_desc(a_lo + ko);
        const uint64_t db_hi = make_sw128_desc(w_hi + ko), db_lo = make_sw128_desc(w_lo + ko);
        const uint32_t first = (it > 0 || kk > 0) ? 1u : 0u;
        tc_mma_tf32(tmem_d, da_hi, db_hi, idesc, first);
        tc_mma_tf32(tmem_d + acc_cols, da_lo, db_hi, idesc, first);
        tc_mma_tf32(tmem_d + acc_cols, da_hi, db_lo, idesc, 1u);
      }
      tc_commit(smem_u32(&sh->empty[s]));
      c.next(kc_n);
      if (++s == P.stages) {
        s = 0;
        ph ^= 1u;
      }
    }
    if (n_active > 0) tc_commit(smem_u32(&sh->accum));
  }
  __syncthreads();
  if (warp == kTcProducers / 32) {
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
  const bool a_vec = ((uintptr_t)a % 16 == 0) && (a_stride % 4 == 0);
  P.out_vec = ((uintptr_t)out % 16 == 0) && (out_stride % 4 == 0);
  {
    const char* dbg = getenv("FSFB_GEMM_DEBUG");
    P.debug = dbg ? atoi(dbg) : 0;
  }
  const int n_w_max = P.S.n_w(0);
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
