// gemm_ss.cu — persistent gather-GEMM with BOTH operands in shared memory (tcgen05.mma "SS" form), fp16-split
// arithmetic.  Same contract as gemm_ts.cu / gemm_tc.cu: out[r] = epi(sum_k a[nbr[k][r]] @ w[k]^T).
//
// Why a third kernel (round-2 measurements, profiles/r2_*): the A-through-TMEM kernel (gemm_ts.cu) gathers with
// thread = row (32 different 128-byte lines per warp instruction).  tools/microbench/gather_paths measures that access
// shape at 22 B/clk/SM against 42-50 B/clk/SM for eight lanes per 128-byte row segment, and switching its MMAs to
// kind::f16 (half the tensor work, half the W bytes) moved the 300 k-point frame by 2 %: the producers' round trip, not the
// tensor pipe, bounds it.  Here
//   * warps 0-15  A producers, all sixteen on every stage: 8 lanes per gathered row read one 128-byte K chunk (four rows per
//     warp instruction, every sector fully used), up to three stages of loads in flight per thread in registers; fp32 →
//     fp16 hi + fp16 (residual * 2048) in registers, lane pairs exchange halves so that every lane writes one 16-byte
//     piece of the SWIZZLE_128B K-major row [hi 64 B | lo 64 B] (conflict-free 128-bit shared stores);
//   * warps 20,21 MMA issuers: per stage two K = 16 steps x (hi*hi → main accumulator, first warp; lo*hi + hi*lo → correction
//     accumulator, second warp, scaled back by 2^-11 in the epilogue) = 6 kind::f16 MMAs instead of the 12 kind::tf32 ones;
//   * warp 22     W loader: one bulk-async copy per stage (fp16 blocks of fsfb_gemm_prepack: 128 B per output channel);
//     slots that already hold the wanted block are not re-copied (Linear layers keep W resident);
//   * warp 23     scheduler: unit ids from a global ticket counter, the unit's 27 x 128 neighbour table (4-byte
//     cp.async) and its active-offset mask, published through an 8-deep info ring;
//   * warps 16-19 epilogue: accumulators are DOUBLE-BUFFERED in tensor memory (2 x (main + correction) x <= 128
//     columns), so the drain of unit i overlaps the MMAs of unit i+1 (the Linear layers of the TS kernel were epilogue
//     bound: one accumulator pair).  Thread = TMEM lane = row, fused bias / LayerNorm / affine / residual / activation,
//     finished rows leave through one bulk-async copy per row.
// A and W rings are independent (a_full/a_empty, w_full/w_empty); the neighbour tables are read by the producers only
// (double-buffered, tbl_empty), unit ids and masks by everyone (info_full/info_empty).  Producers never block on the
// info ring while they hold prefetched stages (mbarrier.test_wait + pipeline bubbles), which is what makes the rings
// deadlock-free when units are shorter than the prefetch depth.
//
// Inputs must stay inside fp16 range (|a| < 65504): the producers track the largest converted magnitude and a launch
// that saw an overflow bumps a device counter (fsfb_gemm_f16_overflows) — results are then Inf/NaN, never silently wrong.
#include <cstdlib>

#include "gemm_persist.cuh"

namespace fsfb {

constexpr int kSsProducerWarps = 16;
constexpr int kSsEpiWarp = kSsProducerWarps;   // warps 16-19 (TMEM lane quarter = warp % 4)
constexpr int kSsMmaWarp = kSsEpiWarp + 4;     // warps 20, 21: MMA issuers (main / correction accumulator)
constexpr int kSsLoaderWarp = kSsMmaWarp + 2;  // warp 22
constexpr int kSsSchedWarp = kSsMmaWarp + 3;   // warp 23
constexpr int kSsThreads = 768;
constexpr int kSsInfo = 8;                     // info ring depth (> prefetch depth + accumulator buffers: see header)
constexpr int kSsMaxA = 6, kSsMaxW = 4;
constexpr uint32_t kSsASlot = kTcRows * 128;   // 128 rows x [hi 64 B | lo 64 B]
constexpr int kSsPrefetch = 4;                 // stages of gathers in flight per producer thread
constexpr int kSsSplitWarps = 4;                 // producer warps of the cp.async pre-split mode (LOAD == 3)
// producer warps / items per thread and stage / register-ring depth of each load mode (an item = 16 bytes of one row's K chunk)
__host__ __device__ constexpr int ss_prod_warps(int load) { return load == 3 ? kSsSplitWarps : (load == 4 ? 8 : kSsProducerWarps); }
__host__ __device__ constexpr int ss_items(int load) { return 1024 / (32 * ss_prod_warps(load)); }

struct SsShared {
  uint64_t a_full[kSsMaxA], a_empty[kSsMaxA];
  uint64_t w_full[kSsMaxW], w_empty[kSsMaxW];
  uint64_t info_full[kSsInfo], info_empty[kSsInfo];
  uint64_t tbl_empty[2];
  uint64_t acc_full[2], acc_empty[2];
  uint64_t res_full[4];
  uint32_t tmem_base;
  uint32_t info_unit[kSsInfo], info_mask[kSsInfo];
};

constexpr int kSsSchedSlots = 256;
__device__ unsigned int g_ss_sched[kSsSchedSlots];
__device__ unsigned int g_ss_overflow;

__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// unit → (row tile, column tile, offset split); column tiles are P.ss_tile_w wide
__device__ __forceinline__ bool ss_unit(const TcParams& P, uint32_t u, TsUnit& U) {
  if (u >= (uint32_t)P.n_units) return false;
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
  U.n_sub = min(P.ss_tile_w, P.S.n_pad() - U.ct * P.ss_tile_w);
  U.k_keep = 0xffffffffu;
  if (P.splits > 1) {
    const uint32_t k_lo = ((uint32_t)P.koff * sp) / (uint32_t)P.splits, k_hi = ((uint32_t)P.koff * (sp + 1)) / (uint32_t)P.splits;
    U.k_keep = (k_hi >= 32 ? 0xffffffffu : ((1u << k_hi) - 1u)) & ~((1u << k_lo) - 1u);
  }
  return true;
}

// LOAD: 0 = complete K chunks from 16-byte aligned rows (cin % 32 == 0): one predicated 128-bit load per item;
//       1 = 16-byte aligned rows, any cin (tail columns masked after the load); 2 = scalar loads (unaligned rows);
//       3 = the input is already in the fp16-split row format of fsfb_split_rows ([rows][cin / 32][hi 32 halves | lo 32
//           halves], i.e. the bytes of a shared-memory operand row): the gather is pure data movement — 16-byte cp.async
//           copies (zero-filled for missing neighbours) that land in the swizzled slot and complete on the slot's mbarrier;
//           four producer warps, no registers, no conversion (the issue slots of the other modes' converts are what bound
//           the 27-offset convolutions: profiles/r2_ss_ncu.md)
// HV: per-channel epilogue vectors in the kernel parameters (cout <= 128, host copies given)
template <int LOAD, bool HV>
__global__ void __launch_bounds__(kSsThreads, 1) k_gather_gemm_ss(const __grid_constant__ TcParams P) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t base = smem_u32(smem_raw);
  const uint32_t s_w = base + P.ss_off_w;
  const uint32_t s_nbr = base + P.ss_off_nbr;
  const uint32_t nbr_bytes = (uint32_t)P.koff * kTcRows * 4u;
  const uint32_t s_stage = base + P.ss_off_stage;
  const uint32_t s_vec = base + P.ss_off_vec;  // [3][256] f32: bias, norm_w, norm_b of the column tile
  SsShared* sh = reinterpret_cast<SsShared*>(smem_raw + P.ss_off_sh);
  const int a_stages = P.ss_a_stages, w_stages = P.ss_w_stages;
  const uint32_t stage_stride = (uint32_t)P.ss_stage_stride;

  if (tid == 0) {
    if (base & 1023u) __trap();
    for (int s = 0; s < kSsMaxA; ++s) {
      mbar_init(smem_u32(&sh->a_full[s]), LOAD == 3 ? kSsSplitWarps * 32 : ss_prod_warps(LOAD));
      mbar_init(smem_u32(&sh->a_empty[s]), 2);  // both MMA issuers commit
    }
    for (int s = 0; s < kSsMaxW; ++s) {
      mbar_init(smem_u32(&sh->w_full[s]), 1);
      mbar_init(smem_u32(&sh->w_empty[s]), 2);
    }
    for (int s = 0; s < kSsInfo; ++s) {
      mbar_init(smem_u32(&sh->info_full[s]), 1);
      mbar_init(smem_u32(&sh->info_empty[s]), ss_prod_warps(LOAD) + 2 + 1 + 4);  // + MMA warps, loader, epilogue
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&sh->tbl_empty[s]), ss_prod_warps(LOAD));
      mbar_init(smem_u32(&sh->acc_full[s]), 2);
      mbar_init(smem_u32(&sh->acc_empty[s]), 4);
    }
    for (int w = 0; w < 4; ++w) mbar_init(smem_u32(&sh->res_full[w]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  constexpr uint32_t tmem_cols = 512;
  if (warp == kSsMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sh->tmem_base)), "r"(tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = sh->tmem_base;
  const uint32_t acc_cols = (uint32_t)P.ss_acc_cols;
  const int kc_n = P.S.kc();
  // role cycle counters (FSFB_GEMM_TIMERS=1, tools/gemm_role_timers.py): phases partition each role's time
  const bool timed = P.timers != nullptr;
  uint32_t t_last = 0;
#define SS_T0() do { if (timed) t_last = (uint32_t)clock(); } while (0)
#define SS_ACC(var) do { if (timed) { const uint32_t t1_ = (uint32_t)clock(); var += t1_ - t_last; t_last = t1_; } } while (0)

  // blocking open of info entry `seq` (roles that hold nothing the others wait for): unit id + mask, entry released
  auto open_info = [&](uint32_t seq, uint32_t& u, uint32_t& m) {
    const uint32_t i8 = seq & (kSsInfo - 1);
    if (lane == 0) mbar_wait(smem_u32(&sh->info_full[i8]), (seq / kSsInfo) & 1u);
    __syncwarp();
    u = lds_u32(smem_u32(&sh->info_unit[i8]));
    m = lds_u32(smem_u32(&sh->info_mask[i8]));
    __syncwarp();
    if (lane == 0) mbar_arrive(smem_u32(&sh->info_empty[i8]));
  };

  if (warp < kSsProducerWarps) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 88;" ::: "memory");
    if (warp < ss_prod_warps(LOAD)) {
    // ================= A producers =================
    const int chunk = tid & 7;    // 16-byte piece (4 floats) of the 128-byte K chunk
    constexpr int kItems = ss_items(LOAD);               // 2 (sixteen warps) or 4 (eight warps: LOAD == 4)
    constexpr int kRowStep = kTcRows / kItems;           // this thread's rows: row_a + kRowStep * p
    constexpr int kRing = kItems == 2 ? kSsPrefetch : 3;  // register-ring depth: 8 / 12 float4 per thread
    const int row_a = tid >> 3;
    const bool odd = (chunk & 1) != 0;
    // 16-byte piece of the fp16 row this lane writes: converted modes pair lanes (even lane: both hi halves, odd: both lo);
    // the pre-split copy (LOAD == 4) moves piece `chunk` as it is
    const uint32_t piece = LOAD == 4 ? (uint32_t)chunk : (uint32_t)(odd ? 4 + (chunk >> 1) : (chunk >> 1));
    uint32_t dst_off[kItems];
#pragma unroll
    for (int p = 0; p < kItems; ++p) {
      const int row = row_a + kRowStep * p;
      dst_off[p] = (uint32_t)row * 128u + ((piece ^ (uint32_t)(row & 7)) << 4);
    }
    // ---- load-side iterator over (unit, active offset, K chunk) ----
    uint32_t it_seq = 0, rem = 0;
    int it_kc = 0;
    int64_t it_row0 = 0;
    uint32_t it_tbl = 0;
    bool ended = false, hold_tbl = false;
    uint32_t held = 0;
    // returns 1 and the stage (k, kc, row0, table) / 0 if the next unit is not published yet (only when !block) / 2 at the end
    auto advance = [&](bool block, int& k, int& kc, int64_t& row0, uint32_t& tbl) -> int {
      if (rem == 0) {
        if (hold_tbl) {  // every lane's table reads of the finished unit fed loads that have been issued
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&sh->tbl_empty[held]));
          hold_tbl = false;
        }
        for (;;) {
          if (ended) return 2;
          const uint32_t i8 = it_seq & (kSsInfo - 1), par = (it_seq / kSsInfo) & 1u;
          if (block) {
            if (lane == 0) mbar_wait(smem_u32(&sh->info_full[i8]), par);
            __syncwarp();
          } else {
            uint32_t ok = 0;
            if (lane == 0) ok = mbar_test(smem_u32(&sh->info_full[i8]), par) ? 1u : 0u;
            ok = __shfl_sync(0xffffffffu, ok, 0);
            if (!ok) return 0;
          }
          const uint32_t u = lds_u32(smem_u32(&sh->info_unit[i8]));
          uint32_t m = lds_u32(smem_u32(&sh->info_mask[i8]));
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&sh->info_empty[i8]));
          TsUnit U;
          if (!ss_unit(P, u, U)) {
            ended = true;
            return 2;
          }
          m &= U.k_keep;
          const uint32_t tb = it_seq & 1u;
          ++it_seq;
          if (m == 0) {  // nothing to gather for this unit
            if (P.nbr && lane == 0) mbar_arrive(smem_u32(&sh->tbl_empty[tb]));
            continue;
          }
          rem = m;
          it_kc = 0;
          it_row0 = U.row0;
          it_tbl = s_nbr + tb * nbr_bytes;
          if (P.nbr) {
            hold_tbl = true;
            held = tb;
          }
          break;
        }
      }
      k = __ffs(rem) - 1;
      kc = it_kc;
      row0 = it_row0;
      tbl = it_tbl;
      if (++it_kc == kc_n) {
        it_kc = 0;
        rem &= rem - 1;
      }
      return 1;
    };
    if constexpr (LOAD == 3) {
      // ---- pre-split input: 128 threads, thread = (16-byte piece, rows r0 + 16 i); row bytes in global = row bytes of the slot
      const int r0 = tid >> 3;
      const uint32_t dst0 = (uint32_t)r0 * 128u + (((uint32_t)chunk ^ (uint32_t)(r0 & 7)) << 4);  // + 2048 i: same row & 7
      const unsigned char* a_bytes = reinterpret_cast<const unsigned char*>(P.a);
      const int64_t row_bytes = P.a_stride * 4;
      uint32_t a_s = 0, a_ph = 0;
      uint32_t tm_empty = 0, tm_fetch = 0, tm_n = 0;
      const uint32_t tm_start = timed ? (uint32_t)clock() : 0u;
      SS_T0();
      for (;;) {
        int k, kc;
        int64_t row0;
        uint32_t tbl;
        if (advance(true, k, kc, row0, tbl) != 1) break;  // blocking is safe: issued copies complete on their own
        if (lane == 0) mbar_wait(smem_u32(&sh->a_empty[a_s]), a_ph ^ 1u);
        __syncwarp();
        SS_ACC(tm_empty);
        const uint32_t dst = base + a_s * kSsASlot + dst0;
        const unsigned char* g0 = a_bytes + (size_t)kc * 128u + (size_t)chunk * 16u;
        int32_t src[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int row = r0 + 16 * i;
          if (P.nbr) {
            src[i] = lds_i32(tbl + (uint32_t)(k * kTcRows + row) * 4u);
          } else {
            const int64_t r = row0 + row;
            src[i] = r < P.rows ? (int32_t)r : -1;
          }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const bool ok = (uint32_t)src[i] < (uint32_t)P.a_rows && !(P.debug & 1);
          const unsigned char* g = g0 + (int64_t)(ok ? src[i] : 0) * row_bytes;
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + 2048u * i), "l"(g), "r"(ok ? 16 : 0) : "memory");
        }
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&sh->a_full[a_s])) : "memory");
        if (++a_s == (uint32_t)a_stages) {
          a_s = 0;
          a_ph ^= 1u;
        }
        SS_ACC(tm_fetch);
        ++tm_n;
      }
      if (timed && tid == 0) {
        uint32_t* t = P.timers + (size_t)blockIdx.x * 32;
        t[0] = tm_empty; t[1] = 0; t[2] = tm_fetch; t[3] = tm_n;
        t[31] = (uint32_t)clock() - tm_start;
      }
    } else {
    // one prefetched stage: kItems items (rows row_a + kRowStep * p) and its K chunk (-1: empty slot)
    struct Pref {
      float4 v[kItems];
      int kc;
    };
    auto load_stage = [&](Pref& q, int k, int kc, int64_t row0, uint32_t tbl) {
      q.kc = kc;
      const int col = kc * kGemmKChunk + chunk * 4;
#pragma unroll
      for (int p = 0; p < kItems; ++p) {
        const int row = row_a + kRowStep * p;
        int32_t src;
        if (P.nbr) {
          src = lds_i32(tbl + (uint32_t)(k * kTcRows + row) * 4u);
        } else {
          const int64_t r = row0 + row;
          src = r < P.rows ? (int32_t)r : -1;
        }
        if ((uint32_t)src >= (uint32_t)P.a_rows || (P.debug & 1)) src = -1;  // also catches negative entries
        const float* g = P.a + (int64_t)(src >= 0 ? src : 0) * P.a_stride + col;
        if (LOAD == 0 || LOAD == 4) {   // (LOAD == 4: the row is fp16-split bytes of the same size and chunking)
          q.v[p] = ldg_pred_f4_na(g, src >= 0);
        } else if (LOAD == 1) {
          q.v[p] = ldg_pred_f4_na(g, src >= 0 && col < P.cin);  // the row's stride covers round_up(cin, 4)
        } else {
          const bool ok = src >= 0;
          q.v[p].x = ldg_pred_f1(g, ok && col < P.cin);
          q.v[p].y = ldg_pred_f1(g + 1, ok && col + 1 < P.cin);
          q.v[p].z = ldg_pred_f1(g + 2, ok && col + 2 < P.cin);
          q.v[p].w = ldg_pred_f1(g + 3, ok && col + 3 < P.cin);
        }
      }
    };
    __half2 ovf = __floats2half2_rn(0.f, 0.f);
    uint32_t a_s = 0, a_ph = 0;
    uint32_t tm_empty = 0, tm_conv = 0, tm_fetch = 0, tm_n = 0;
    const uint32_t tm_start = timed ? (uint32_t)clock() : 0u;
    auto store_stage = [&](Pref& q) {
      if (lane == 0) mbar_wait(smem_u32(&sh->a_empty[a_s]), a_ph ^ 1u);
      __syncwarp();
      SS_ACC(tm_empty);
      const uint32_t slot = base + a_s * kSsASlot;
#pragma unroll
      for (int p = 0; p < kItems; ++p) {
        float4 v = q.v[p];
        if constexpr (LOAD == 4) {   // pre-split rows: a plain copy into the swizzled slot
          if (!(P.debug & 8)) sts_f4(slot + dst_off[p], v);
          continue;
        }
        if (LOAD == 1) {  // columns past cin inside the last 128-bit piece hold whatever follows the row: zero them
          const int nv = P.cin - (q.kc * kGemmKChunk + chunk * 4);
          if (nv < 4) {
            if (nv < 2) v.y = 0.f;
            if (nv < 3) v.z = 0.f;
            v.w = 0.f;
            if (nv < 1) v.x = 0.f;
          }
        }
        const __half2 h01 = __floats2half2_rn(v.x, v.y), h23 = __floats2half2_rn(v.z, v.w);
        ovf = __hmax2(ovf, __hmax2(__habs2(h01), __habs2(h23)));
        const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
        const __half2 l01 = __floats2half2_rn((v.x - f01.x) * kF16LoScale, (v.y - f01.y) * kF16LoScale);
        const __half2 l23 = __floats2half2_rn((v.z - f23.x) * kF16LoScale, (v.w - f23.y) * kF16LoScale);
        const uint32_t hi0 = *reinterpret_cast<const uint32_t*>(&h01), hi1 = *reinterpret_cast<const uint32_t*>(&h23);
        const uint32_t lo0 = *reinterpret_cast<const uint32_t*>(&l01), lo1 = *reinterpret_cast<const uint32_t*>(&l23);
        // lane pair (even, odd) = 8 consecutive inputs: the even lane keeps both hi halves, the odd lane both lo halves
        const uint32_t r0 = __shfl_xor_sync(0xffffffffu, odd ? hi0 : lo0, 1);
        const uint32_t r1 = __shfl_xor_sync(0xffffffffu, odd ? hi1 : lo1, 1);
        if (!(P.debug & 8)) {
          if (odd) sts_u4(slot + dst_off[p], r0, r1, lo0, lo1);
          else sts_u4(slot + dst_off[p], hi0, hi1, r0, r1);
        }
      }
      if (!(P.debug & 16)) fence_proxy_async();  // generic-proxy stores before the tensor core's async-proxy reads
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&sh->a_full[a_s]));
      if (++a_s == (uint32_t)a_stages) {
        a_s = 0;
        a_ph ^= 1u;
      }
      SS_ACC(tm_conv);
      ++tm_n;
    };
    // Register ring of kSsPrefetch stages with STATIC slot indices (the loop body is unrolled over the slots: no register
    // moves, so nothing waits for a load before its own store).  Visit d: store the stage slot d holds (fetched
    // kSsPrefetch visits ago), then fetch the next stage into it; stages are therefore stored in fetch order whatever
    // bubbles the non-blocking fetches leave.  A fetch blocks only while no slot holds a stage.
    Pref q[kRing];
#pragma unroll
    for (int d = 0; d < kRing; ++d) q[d].kc = -1;
    int n_held = 0;
    SS_T0();
    while (!(ended && n_held == 0)) {
#pragma unroll
      for (int d = 0; d < kRing; ++d) {
        if (q[d].kc >= 0) {
          store_stage(q[d]);
          q[d].kc = -1;
          --n_held;
        }
        if (!ended) {
          int k, kc;
          int64_t row0;
          uint32_t tbl;
          if (advance(n_held == 0, k, kc, row0, tbl) == 1) {
            load_stage(q[d], k, kc, row0, tbl);
            ++n_held;
          }
          SS_ACC(tm_fetch);
        }
      }
    }
    if (timed && tid == 0) {
      uint32_t* t = P.timers + (size_t)blockIdx.x * 32;
      t[0] = tm_empty; t[1] = tm_conv; t[2] = tm_fetch; t[3] = tm_n;
      t[31] = (uint32_t)clock() - tm_start;
    }
    // overflow report: any converted magnitude that became +Inf
    {
      const __half2 inf2 = __floats2half2_rn(65504.f, 65504.f);
      const __half2 gt = __hgt2(ovf, inf2);  // Inf > 65504
      const bool bad = __low2float(gt) != 0.f || __high2float(gt) != 0.f;
      if (__any_sync(0xffffffffu, bad) && lane == 0 && P.ss_overflow) atomicAdd(P.ss_overflow, 1u);
    }
    }  // LOAD != 3
    }  // producing warp
  } else if (warp >= kSsMmaWarp) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;" ::: "memory");
    if (warp == kSsMmaWarp || warp == kSsMmaWarp + 1) {
      // ================= two MMA issuers =================
      // first warp: main accumulator (a_hi * w_hi); second warp: correction accumulator (a_lo * w_hi + a_hi * w_lo).  Two warps
      // because the per-stage instruction latency of ONE issuing warp (waits, descriptors, 6 MMAs, 3 commits: ~650 clk measured)
      // exceeded the 384 clk of tensor work of a stage.  Whole-warp loop with warp-uniform operands, one elected lane issues.
      const bool is_main = warp == kSsMmaWarp;
      const uint32_t u_tmem_d = __shfl_sync(0xffffffffu, tmem_d, 0);
      // low words of the SWIZZLE_128B K-major descriptors of slot 0 (high word constant: SBO = 1024, version 1, layout 2)
      const uint32_t a_d0 = __shfl_sync(0xffffffffu, ((base & 0x3FFFFu) >> 4) | 0x10000u, 0);
      const uint32_t w_d0 = __shfl_sync(0xffffffffu, ((s_w & 0x3FFFFu) >> 4) | 0x10000u, 0);
      const uint32_t w_dstep = __shfl_sync(0xffffffffu, P.ss_w_slot >> 4, 0);
      constexpr uint64_t kDescHi = (uint64_t)0x40004040u << 32;
      const uint32_t a_full0 = __shfl_sync(0xffffffffu, smem_u32(&sh->a_full[0]), 0);
      const uint32_t a_empty0 = __shfl_sync(0xffffffffu, smem_u32(&sh->a_empty[0]), 0);
      const uint32_t w_full0 = __shfl_sync(0xffffffffu, smem_u32(&sh->w_full[0]), 0);
      const uint32_t w_empty0 = __shfl_sync(0xffffffffu, smem_u32(&sh->w_empty[0]), 0);
      int a_s = 0, w_s = 0;
      uint32_t a_ph = 0, w_ph = 0;
      uint32_t tm_open = 0, tm_w = 0, tm_a = 0, tm_issue = 0;
      SS_T0();
      for (uint32_t seq = 0;; ++seq) {
        uint32_t u, m0;
        open_info(seq, u, m0);
        TsUnit U;
        if (!ss_unit(P, __shfl_sync(0xffffffffu, u, 0), U)) break;
        const uint32_t m = __shfl_sync(0xffffffffu, m0, 0) & U.k_keep;
        const int n_active = __popc(m) * kc_n;
        const uint32_t idesc = make_idesc_f16(U.n_sub);
        const uint32_t buf = P.ss_bufs == 2 ? (seq & 1u) : 0u;
        const uint32_t use = P.ss_bufs == 2 ? (seq >> 1) : seq;  // earlier uses of this accumulator buffer
        const uint32_t d_acc = u_tmem_d + buf * 2u * acc_cols + (is_main ? 0u : acc_cols);
        const uint32_t acc_full_bar = smem_u32(&sh->acc_full[buf]);
        if (use > 0) {  // the epilogue has read the buffer's previous unit out of TMEM
          if (lane == 0) mbar_wait(smem_u32(&sh->acc_empty[buf]), (use - 1u) & 1u);
          __syncwarp();
          tc_fence_after();
        }
        SS_ACC(tm_open);
        int kc = 0;
        for (int it = 0; it < n_active; ++it) {
          if (lane == 0) mbar_wait(w_full0 + 8u * w_s, w_ph);
          SS_ACC(tm_w);
          if (lane == 0) mbar_wait(a_full0 + 8u * a_s, a_ph);
          __syncwarp();
          tc_fence_after();
          SS_ACC(tm_a);
          const uint32_t a_d = a_d0 + (uint32_t)a_s * (kSsASlot >> 4);
          const uint32_t w_d = w_d0 + (uint32_t)w_s * w_dstep;
          int ksteps = 2;
          if (LOAD != 0 && LOAD != 3 && LOAD != 4) ksteps = (min(kGemmKChunk, P.cin - kc * kGemmKChunk) + 15) >> 4;
          uint32_t elected;
          asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(elected));
          if (elected) {
            if (!(P.debug & 4)) {
              // 16-byte units inside the 128-byte row: K step +2, lo half +4
#pragma unroll
              for (int kk = 0; kk < 2; ++kk) {
                if (kk < ksteps) {
                  const uint32_t acc = (it > 0 || kk > 0) ? 1u : 0u;
                  if (is_main) {
                    tc_mma_f16_ss(d_acc, kDescHi | (a_d + 2u * kk), kDescHi | (w_d + 2u * kk), idesc, acc);
                  } else {
                    tc_mma_f16_ss(d_acc, kDescHi | (a_d + 4u + 2u * kk), kDescHi | (w_d + 2u * kk), idesc, acc);
                    tc_mma_f16_ss(d_acc, kDescHi | (a_d + 2u * kk), kDescHi | (w_d + 4u + 2u * kk), idesc, 1u);
                  }
                }
              }
            }
            tc_commit(a_empty0 + 8u * a_s);
            tc_commit(w_empty0 + 8u * w_s);
            if (it == n_active - 1) tc_commit(acc_full_bar);
          }
          __syncwarp();
          if (++kc == kc_n) kc = 0;
          if (++a_s == a_stages) {
            a_s = 0;
            a_ph ^= 1u;
          }
          if (++w_s == w_stages) {
            w_s = 0;
            w_ph ^= 1u;
          }
          SS_ACC(tm_issue);
        }
        if (n_active == 0 && lane == 0) mbar_arrive(acc_full_bar);  // the epilogue writes zeros
      }
      if (timed && lane == 0 && is_main) {
        uint32_t* t = P.timers + (size_t)blockIdx.x * 32 + 4;
        t[0] = tm_open; t[1] = tm_w; t[2] = tm_a; t[3] = tm_issue;
      }
    } else if (warp == kSsLoaderWarp) {
      // ================= W loader =================
      int w_s = 0;
      uint32_t w_ph = 0;
      int t0 = -1, t1 = -1, t2 = -1, t3 = -1;  // block held by each slot (registers: no dynamically indexed array)
      for (uint32_t seq = 0;; ++seq) {
        uint32_t u, m;
        open_info(seq, u, m);
        TsUnit U;
        if (!ss_unit(P, u, U)) break;
        m &= U.k_keep;
        const int n_active = __popc(m) * kc_n;
        const int c0 = U.ct * P.ss_tile_w;
        const int nt256 = c0 / kGemmNTile;
        const size_t blk_bytes = P.S.f16_block_bytes(nt256);
        const unsigned char* w_unit = P.w_packed + P.S.f16_tile_base(nt256) + (size_t)(c0 % kGemmNTile) * 128u;
        const uint32_t sub_bytes = (uint32_t)U.n_sub * 128u;
        const int tag_base = U.ct * P.koff * kc_n;
        StageCursor c;
        c.init(m);
        for (int it = 0; it < n_active; ++it) {
          const int blk = c.k * kc_n + c.kc;
          const int want = tag_base + blk;
          const int have = w_s == 0 ? t0 : (w_s == 1 ? t1 : (w_s == 2 ? t2 : t3));
          if (lane == 0) {
            mbar_wait(smem_u32(&sh->w_empty[w_s]), w_ph ^ 1u);
            if (have != want && !(P.debug & 2)) {
              mbar_expect_tx(smem_u32(&sh->w_full[w_s]), sub_bytes);
              bulk_g2s(s_w + (uint32_t)w_s * P.ss_w_slot, w_unit + (size_t)blk * blk_bytes, sub_bytes, smem_u32(&sh->w_full[w_s]));
            }
            mbar_arrive(smem_u32(&sh->w_full[w_s]));
          }
          if (w_s == 0) t0 = want; else if (w_s == 1) t1 = want; else if (w_s == 2) t2 = want; else t3 = want;
          __syncwarp();
          c.next(kc_n);
          if (++w_s == w_stages) {
            w_s = 0;
            w_ph ^= 1u;
          }
        }
      }
    } else if (warp == kSsSchedWarp) {
      // ================= scheduler =================
      for (uint32_t seq = 0;; ++seq) {
        const uint32_t i8 = seq & (kSsInfo - 1);
        if (seq >= (uint32_t)kSsInfo) {
          if (lane == 0) mbar_wait(smem_u32(&sh->info_empty[i8]), ((seq / kSsInfo) + 1u) & 1u);
          __syncwarp();
        }
        uint32_t u = 0;
        if (lane == 0) u = atomicAdd(&g_ss_sched[P.sched_slot], 1u) - P.sched_base;
        u = __shfl_sync(0xffffffffu, u, 0);
        TsUnit U;
        const bool have = ss_unit(P, u, U);
        uint32_t my_mask = have ? 1u : 0u;  // Linear: the single "offset"
        if (have && P.nbr) {
          const uint32_t tb = seq & 1u;
          const uint32_t nb = s_nbr + tb * nbr_bytes;
          if (seq >= 2u) {  // the producers have left the unit that used this table buffer
            if (lane == 0) mbar_wait(smem_u32(&sh->tbl_empty[tb]), ((seq >> 1) + 1u) & 1u);
            __syncwarp();
          }
          if (P.nbr_ro) {   // table permuted into the row order, padded to whole tiles: 27 contiguous 512-byte runs
            const int64_t stride = (P.rows + kTcRows - 1) / kTcRows * kTcRows;
            const int32_t* src = P.nbr + U.row0 + 4 * lane;
            for (int k = 0; k < P.koff; ++k)
              asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(nb + (uint32_t)(k * kTcRows + 4 * lane) * 4u),
                           "l"(src + (int64_t)k * stride)
                           : "memory");
          } else {
          int64_t r[4];
#pragma unroll
          for (int rq = 0; rq < 4; ++rq) {
            const int r_l = 32 * rq + lane;
            r[rq] = -1;
            if (U.row0 + r_l < P.rows) r[rq] = P.row_order ? (int64_t)__ldg(P.row_order + U.row0 + r_l) : U.row0 + r_l;
          }
#pragma unroll
          for (int rq = 0; rq < 4; ++rq) {
            const int r_l = 32 * rq + lane;
            if (r[rq] >= 0) {
              for (int k = 0; k < P.koff; ++k)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(nb + (uint32_t)(k * kTcRows + r_l) * 4u),
                             "l"(P.nbr + (int64_t)k * P.rows + r[rq])
                             : "memory");
            } else {
              for (int k = 0; k < P.koff; ++k) sts_i32(nb + (uint32_t)(k * kTcRows + r_l) * 4u, -1);
            }
          }
          }
          asm volatile("cp.async.wait_all;" ::: "memory");
          __syncwarp();
          my_mask = 0;
          for (int k = 0; k < P.koff; ++k) {
            bool any = false;
#pragma unroll
            for (int rq = 0; rq < 4; ++rq) {
              const int32_t src = lds_i32(nb + (uint32_t)(k * kTcRows + 32 * rq + lane) * 4u);
              any |= (uint32_t)src < (uint32_t)P.a_rows;
            }
            if (any) my_mask |= 1u << k;
          }
          my_mask = __reduce_or_sync(0xffffffffu, my_mask);
        }
        if (lane == 0) {
          sh->info_unit[i8] = u;
          sh->info_mask[i8] = my_mask;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&sh->info_full[i8]));
        if (!have) break;
      }
    }
  } else {
    // ================= epilogue warps (TMEM lane quarter q = warp % 4) =================
    const int q = warp & 3;
    const int r_l = 32 * q + lane;
    const uint32_t lane_off = (uint32_t)(32 * q) << 16;
    const uint32_t my_row = s_stage + (uint32_t)r_l * stage_stride * 4u;
    const uint32_t my_stage = s_stage + (uint32_t)(32 * q) * stage_stride * 4u;
    const uint32_t res_bar = smem_u32(&sh->res_full[q]);
    const Epilogue& E = P.E;
    const int act = E.act & 0xff;
    const bool post = (E.act & FSFB_RESIDUAL_POST) != 0;
    const bool res_vec = !E.residual || (((uintptr_t)E.residual % 16 == 0) && (E.residual_stride % 4 == 0));
    uint32_t res_ph = 0;
    int vec_ct = -1;
    uint32_t tm_pre = 0, tm_accf = 0, tm_drain = 0, tm_out = 0, tm_units = 0;
    SS_T0();
    for (uint32_t seq = 0;; ++seq) {
      uint32_t u, m;
      open_info(seq, u, m);
      TsUnit U;
      if (!ss_unit(P, u, U)) break;
      m &= U.k_keep;
      int64_t r_cur = P.rows;
      if (U.row0 + r_l < P.rows) r_cur = P.row_order ? (int64_t)__ldg(P.row_order + U.row0 + r_l) : U.row0 + r_l;
      const uint32_t buf = P.ss_bufs == 2 ? (seq & 1u) : 0u;
      const uint32_t use = P.ss_bufs == 2 ? (seq >> 1) : seq;
      const uint32_t t_row = tmem_d + buf * 2u * acc_cols + lane_off;
      const uint32_t acc_empty_bar = smem_u32(&sh->acc_empty[buf]);
      const bool have_acc = m != 0;
      const int c0 = U.ct * P.ss_tile_w;
      const int c_n = min(U.n_sub, P.S.cout - c0);  // real channels in this column tile
      const bool split = P.splits > 1;
      const bool fast = split || ((c_n & 3) == 0 && P.out_vec && res_vec);
      // odd widths / unaligned rows (33, 11, 131 ... channel heads): the same fused thread-per-row epilogue into the staging
      // tile, then a flat element walk with 4-byte stores that are contiguous within each row
      const bool generic = !fast && E.residual == nullptr;
      const bool fused = (fast && !split) || generic;
      const bool valid = r_cur < P.rows;
      // the previous unit's bulk stores have read this thread's staging row
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      __syncwarp();
      const bool use_res = fused && E.residual != nullptr;
      if (use_res) {  // residual rows → staging rows (bulk async, lands while the MMAs of this unit still run)
        const uint32_t n_valid = __popc(__ballot_sync(0xffffffffu, valid));
        if (lane == 0) mbar_expect_tx(res_bar, n_valid * (uint32_t)c_n * 4u);
        __syncwarp();
        if (valid) bulk_g2s(my_row, E.residual + r_cur * E.residual_stride + c0, (uint32_t)c_n * 4u, res_bar);
        if (lane == 0) mbar_arrive(res_bar);
      }
      if (!HV && fused && vec_ct != U.ct) {  // per-channel vectors of this column tile (missing ones default to no-ops)
        asm volatile("bar.sync 1, 128;" ::: "memory");  // every epilogue warp is done with the previous tile's vectors
        float* sv = reinterpret_cast<float*>(smem_raw + P.ss_off_vec);
        for (int j = r_l; j < 256; j += 128) {
          const int c = c0 + j;
          const bool in = j < c_n;
          sv[j] = (in && E.bias) ? __ldg(E.bias + c) : 0.f;
          sv[256 + j] = (in && E.norm_w) ? __ldg(E.norm_w + c) : 1.f;
          sv[512 + j] = (in && E.norm_b) ? __ldg(E.norm_b + c) : 0.f;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        vec_ct = U.ct;
      }
      SS_ACC(tm_pre);
      if (lane == 0) mbar_wait(smem_u32(&sh->acc_full[buf]), use & 1u);
      __syncwarp();
      tc_fence_after();
      SS_ACC(tm_accf);
      const bool blk32 = split || (HV && fused);
      if (blk32) {
        if (use_res) {
          mbar_wait(res_bar, res_ph);
          res_ph ^= 1u;
        }
        const int nrm = (HV && fused) ? E.norm : FSFB_NORM_NONE;
        const int ac = (HV && fused) ? act : FSFB_ACT_NONE;
        const int cn = (HV && fused) ? c_n : 0;  // raw sums (offset splits): no column is finished here
#define SS_EPI(N, A, PO) ts_epi32<N, A, PO, true, true>(P, t_row, acc_cols, my_row, cn, U.n_sub, have_acc, use_res, valid, acc_empty_bar, lane)
        if (ac == FSFB_ACT_RELU) {
          if (nrm == FSFB_NORM_LAYERNORM) { if (post) SS_EPI(FSFB_NORM_LAYERNORM, FSFB_ACT_RELU, true); else SS_EPI(FSFB_NORM_LAYERNORM, FSFB_ACT_RELU, false); }
          else if (nrm == FSFB_NORM_AFFINE) { if (post) SS_EPI(FSFB_NORM_AFFINE, FSFB_ACT_RELU, true); else SS_EPI(FSFB_NORM_AFFINE, FSFB_ACT_RELU, false); }
          else { if (post) SS_EPI(FSFB_NORM_NONE, FSFB_ACT_RELU, true); else SS_EPI(FSFB_NORM_NONE, FSFB_ACT_RELU, false); }
        } else {
          if (nrm == FSFB_NORM_LAYERNORM) SS_EPI(FSFB_NORM_LAYERNORM, FSFB_ACT_NONE, false);
          else if (nrm == FSFB_NORM_AFFINE) SS_EPI(FSFB_NORM_AFFINE, FSFB_ACT_NONE, false);
          else SS_EPI(FSFB_NORM_NONE, FSFB_ACT_NONE, false);
        }
#undef SS_EPI
      } else if (fused && !use_res) {
        // Vectors from shared memory (GELU layers, wide tiles, device-only vectors), no residual: ONE pass over tensor memory
        // (32 columns per tcgen05.ld pair) that parks the biased sums in the thread's own staging row and hands the
        // accumulators back; LayerNorm statistics and the norm / activation then run from shared memory.  (The three-pass form
        // below re-read TMEM for mean, variance and output, 8 columns per round trip: 40-50 k clk per unit on LN + GELU layers.)
        const bool has_bias = E.bias != nullptr;
        float sum = 0.f;
#pragma unroll 1
        for (int cb = 0; cb < U.n_sub; cb += 32) {
          float v[32];
          if (have_acc) {
            float c2[32];
            tc_ld32(t_row + cb, v);
            tc_ld32(t_row + acc_cols + cb, c2);
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) v[jj] = fmaf(c2[jj], 1.f / kF16LoScale, v[jj]);
          } else {
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) v[jj] = 0.f;
          }
          if (cb + 32 >= U.n_sub) {  // last TMEM read of this unit
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty_bar);
          }
#pragma unroll
          for (int jj = 0; jj < 32; jj += 4) {
            float4 y = make_float4(v[jj], v[jj + 1], v[jj + 2], v[jj + 3]);
            if (has_bias) {
              const float4 bb = lds_f4(s_vec + (uint32_t)(cb + jj) * 4u);
              y.x += bb.x; y.y += bb.y; y.z += bb.z; y.w += bb.w;
            }
            // columns in [c_n, n_sub) are exact zeros (zero weight rows, default bias 0); beyond n_sub the 32-column load
            // returns whatever the tensor memory held
            if (cb + jj < U.n_sub) sum += (y.x + y.y) + (y.z + y.w);
            sts_f4(my_row + (uint32_t)(cb + jj) * 4u, y);
          }
        }
        float mean = 0.f, rstd = 1.f;
        if (E.norm == FSFB_NORM_LAYERNORM) {
          mean = sum / (float)c_n;
          float qq = 0.f;
#pragma unroll 4
          for (int c = 0; c < c_n; c += 4) {
            const float4 x = lds_f4(my_row + (uint32_t)c * 4u);
            const float d0 = x.x - mean, d1 = x.y - mean, d2 = x.z - mean, d3 = x.w - mean;
            qq += d0 * d0 + (c + 1 < c_n ? d1 * d1 : 0.f) + (c + 2 < c_n ? d2 * d2 : 0.f) + (c + 3 < c_n ? d3 * d3 : 0.f);
          }
          rstd = 1.f / sqrtf(qq / (float)c_n + E.eps);
        }
        if (E.norm != FSFB_NORM_NONE || act != FSFB_ACT_NONE) {
#pragma unroll 2
          for (int c = 0; c < c_n; c += 4) {
            float4 y = lds_f4(my_row + (uint32_t)c * 4u);
            if (E.norm != FSFB_NORM_NONE) {
              const float4 w4 = lds_f4(s_vec + (uint32_t)(256 + c) * 4u), h4 = lds_f4(s_vec + (uint32_t)(512 + c) * 4u);
              if (E.norm == FSFB_NORM_LAYERNORM) {
                y.x = (y.x - mean) * rstd * w4.x + h4.x; y.y = (y.y - mean) * rstd * w4.y + h4.y;
                y.z = (y.z - mean) * rstd * w4.z + h4.z; y.w = (y.w - mean) * rstd * w4.w + h4.w;
              } else {
                y.x = fmaf(y.x, w4.x, h4.x); y.y = fmaf(y.y, w4.y, h4.y); y.z = fmaf(y.z, w4.z, h4.z); y.w = fmaf(y.w, w4.w, h4.w);
              }
            }
            y.x = apply_act(y.x, act); y.y = apply_act(y.y, act); y.z = apply_act(y.z, act); y.w = apply_act(y.w, act);
            sts_f4(my_row + (uint32_t)c * 4u, y);
          }
        }
      } else {
        float v[8];
        auto ld_issue = [&](int cb, uint32_t(&a)[8], uint32_t(&b)[8]) {  // warp-collective; pair with ld_wait
          if (have_acc) {
            tc_ld8_nowait(t_row + cb, a);
            tc_ld8_nowait(t_row + acc_cols + cb, b);
          }
        };
        auto ld_wait = [&](const uint32_t(&a)[8], const uint32_t(&b)[8]) {
          if (have_acc) {
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) v[jj] = fmaf(__uint_as_float(b[jj]), 1.f / kF16LoScale, __uint_as_float(a[jj]));
          } else {
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) v[jj] = 0.f;
          }
        };
        const bool has_bias = E.bias != nullptr;
        float mean = 0.f, rstd = 1.f;
        if (fused && E.norm == FSFB_NORM_LAYERNORM) {  // two-pass row statistics: the whole row is in this tile
          float sum = 0.f;
          for (int cb = 0; cb < c_n; cb += 8) {
            uint32_t a[8], b[8];
            ld_issue(cb, a, b);
            ld_wait(a, b);
#pragma unroll
            for (int jj = 0; jj < 8; jj += 4) {   // columns >= c_n hold exact zeros (zero weight rows, default bias 0)
              float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
              if (has_bias) bb = lds_f4(s_vec + (uint32_t)(cb + jj) * 4u);
              if (cb + jj < c_n) sum += (v[jj] + bb.x) + (v[jj + 1] + bb.y) + (v[jj + 2] + bb.z) + (v[jj + 3] + bb.w);
            }
          }
          mean = sum / (float)c_n;
          float qq = 0.f;
          for (int cb = 0; cb < c_n; cb += 8) {
            uint32_t a[8], b[8];
            ld_issue(cb, a, b);
            ld_wait(a, b);
#pragma unroll
            for (int jj = 0; jj < 8; jj += 4) {
              float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
              if (has_bias) bb = lds_f4(s_vec + (uint32_t)(cb + jj) * 4u);
              if (cb + jj < c_n) {   // a width that is not a multiple of 4 ends inside this group: its tail columns do not count
                const float d0 = v[jj] + bb.x - mean, d1 = v[jj + 1] + bb.y - mean, d2 = v[jj + 2] + bb.z - mean, d3 = v[jj + 3] + bb.w - mean;
                qq += d0 * d0 + (cb + jj + 1 < c_n ? d1 * d1 : 0.f) + (cb + jj + 2 < c_n ? d2 * d2 : 0.f) + (cb + jj + 3 < c_n ? d3 * d3 : 0.f);
              }
            }
          }
          rstd = 1.f / sqrtf(qq / (float)c_n + E.eps);
        }
        if (use_res) {
          mbar_wait(res_bar, res_ph);
          res_ph ^= 1u;
        }
        // thread = TMEM lane = row: accumulators (+ fused epilogue in fast mode) → staging row, 8 columns at a time
        for (int cb = 0; cb < U.n_sub; cb += 8) {
          uint32_t a[8], b[8];
          ld_issue(cb, a, b);
          const uint32_t srow = my_row + (uint32_t)cb * 4u;
          const bool live = fused && cb < c_n;  // c_n % 4 == 0 and the tile's padding is whole 8-column groups
          float4 vb4[2], vw4[2], vh4[2], g4[2];
#pragma unroll
          for (int t = 0; t < 2; ++t) {  // requested before the TMEM wait: independent of it
            vb4[t] = make_float4(0.f, 0.f, 0.f, 0.f);
            vw4[t] = make_float4(1.f, 1.f, 1.f, 1.f);
            vh4[t] = make_float4(0.f, 0.f, 0.f, 0.f);
            g4[t] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (live) {
              if (has_bias) vb4[t] = lds_f4(s_vec + (uint32_t)(cb + 4 * t) * 4u);
              if (E.norm != FSFB_NORM_NONE) {
                vw4[t] = lds_f4(s_vec + (uint32_t)(256 + cb + 4 * t) * 4u);
                vh4[t] = lds_f4(s_vec + (uint32_t)(512 + cb + 4 * t) * 4u);
              }
              if (use_res && valid && cb + 4 * t < c_n) g4[t] = lds_f4(srow + 16 * t);
            }
          }
          ld_wait(a, b);
          if (cb + 8 >= U.n_sub) {  // last TMEM read of this unit: the MMA warp may reuse the buffer
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty_bar);
          }
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            float4 y = make_float4(v[4 * t], v[4 * t + 1], v[4 * t + 2], v[4 * t + 3]);
            if (live) {
              y.x += vb4[t].x; y.y += vb4[t].y; y.z += vb4[t].z; y.w += vb4[t].w;
              if (E.norm == FSFB_NORM_LAYERNORM) {
                y.x = (y.x - mean) * rstd * vw4[t].x + vh4[t].x; y.y = (y.y - mean) * rstd * vw4[t].y + vh4[t].y;
                y.z = (y.z - mean) * rstd * vw4[t].z + vh4[t].z; y.w = (y.w - mean) * rstd * vw4[t].w + vh4[t].w;
              } else if (E.norm == FSFB_NORM_AFFINE) {
                y.x = fmaf(y.x, vw4[t].x, vh4[t].x); y.y = fmaf(y.y, vw4[t].y, vh4[t].y);
                y.z = fmaf(y.z, vw4[t].z, vh4[t].z); y.w = fmaf(y.w, vw4[t].w, vh4[t].w);
              }
              if (post) {
                y.x = apply_act(y.x, act) + g4[t].x; y.y = apply_act(y.y, act) + g4[t].y; y.z = apply_act(y.z, act) + g4[t].z; y.w = apply_act(y.w, act) + g4[t].w;
              } else {
                y.x = apply_act(y.x + g4[t].x, act); y.y = apply_act(y.y + g4[t].y, act); y.z = apply_act(y.z + g4[t].z, act); y.w = apply_act(y.w + g4[t].w, act);
              }
            }
            sts_f4(srow + 16 * t, y);
          }
        }
      }
      SS_ACC(tm_drain);
      // ---- rows leave shared memory ----
      if (fast) {
        fence_proxy_async();  // this thread's staging writes (generic proxy) before its bulk copy's reads (async proxy)
        if (valid) {
          float* dst = split ? P.partial + ((int64_t)U.sp * P.rows + r_cur) * P.cpad + c0 : P.out + r_cur * P.out_stride + c0;
          const uint32_t bytes = (uint32_t)(split ? U.n_sub : c_n) * 4u;
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(my_row), "r"(bytes) : "memory");
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      } else if (generic) {
        __syncwarp();  // this warp's 32 staging rows are complete
        const int dq = 32 / c_n, dr = 32 - dq * c_n;  // element e = 32 it + lane of the warp's [32, c_n] block
        int r = lane / c_n, c = lane - r * c_n;
        for (int it = 0; it < c_n; ++it) {  // 32 c_n elements / 32 lanes: every lane has exactly c_n of them
          const int64_t rr = __shfl_sync(0xffffffffu, r_cur, r);
          float x;
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"(my_stage + (uint32_t)(r * (int)stage_stride + c) * 4u));
          if (rr < P.rows) P.out[rr * P.out_stride + c0 + c] = x;
          r += dq;
          c += dr;
          if (c >= c_n) {
            c -= c_n;
            ++r;
          }
        }
        __syncwarp();  // reads of the staging rows finish before the next unit overwrites them
      } else {
        __syncwarp();  // this warp's staging rows are complete
        Epilogue Es = E;
        if (Es.bias) Es.bias += c0;
        if (Es.norm_w) Es.norm_w += c0;
        if (Es.norm_b) Es.norm_b += c0;
        if (Es.residual) Es.residual += c0;
        for (int i = 0; i < 32; ++i) {
          const int64_t r = __shfl_sync(0xffffffffu, r_cur, i);
          if (r >= P.rows) continue;  // warp-uniform
          const float* xs = reinterpret_cast<const float*>(smem_raw + (my_stage - base) + (size_t)i * stage_stride * 4);
          warp_row_epilogue(xs, c_n, Es, r, P.out + r * P.out_stride + c0);
        }
        __syncwarp();  // reads of the staging rows finish before the next unit overwrites them
      }
      SS_ACC(tm_out);
      ++tm_units;
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // the last rows are on their way out before the CTA exits
    if (timed && warp == kSsEpiWarp && lane == 0) {
      uint32_t* t = P.timers + (size_t)blockIdx.x * 32 + 8;
      t[0] = tm_pre; t[1] = tm_accf; t[2] = tm_drain; t[3] = tm_out; t[4] = tm_units;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kSsMmaWarp) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(tmem_cols) : "memory");
  }
}

// fp32 rows → the fp16-split row format the LOAD == 3 producers copy: per 32 input channels one 128-byte block
// [32 x fp16(a) | 32 x fp16((a - hi) * 2048)]; thread = 8 channels (two 16-byte pieces).  HBM-bound: 4 c B read + 4 c B written per row.
__global__ void __launch_bounds__(256) k_split_rows(const float* __restrict__ a, int64_t rows, int c, int64_t a_stride,
                                                    unsigned char* __restrict__ out, unsigned int* overflow) {
  const int per_row = c >> 3;
  const int64_t total = rows * per_row;
  __half2 ovf = __floats2half2_rn(0.f, 0.f);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / per_row;
    const int t = (int)(i - r * per_row);
    const float4* src = reinterpret_cast<const float4*>(a + r * a_stride + 8 * t);
    const float4 x = ldg_stream_f4(src), y = ldg_stream_f4(src + 1);
    const __half2 h0 = __floats2half2_rn(x.x, x.y), h1 = __floats2half2_rn(x.z, x.w), h2 = __floats2half2_rn(y.x, y.y), h3 = __floats2half2_rn(y.z, y.w);
    ovf = __hmax2(ovf, __hmax2(__hmax2(__habs2(h0), __habs2(h1)), __hmax2(__habs2(h2), __habs2(h3))));
    const float2 f0 = __half22float2(h0), f1 = __half22float2(h1), f2 = __half22float2(h2), f3 = __half22float2(h3);
    const __half2 l0 = __floats2half2_rn((x.x - f0.x) * kF16LoScale, (x.y - f0.y) * kF16LoScale);
    const __half2 l1 = __floats2half2_rn((x.z - f1.x) * kF16LoScale, (x.w - f1.y) * kF16LoScale);
    const __half2 l2 = __floats2half2_rn((y.x - f2.x) * kF16LoScale, (y.y - f2.y) * kF16LoScale);
    const __half2 l3 = __floats2half2_rn((y.z - f3.x) * kF16LoScale, (y.w - f3.y) * kF16LoScale);
    unsigned char* dst = out + (r * (int64_t)c) * 4 + (int64_t)(t >> 2) * 128 + (t & 3) * 16;
    uint4 hi, lo;
    hi.x = *reinterpret_cast<const uint32_t*>(&h0); hi.y = *reinterpret_cast<const uint32_t*>(&h1);
    hi.z = *reinterpret_cast<const uint32_t*>(&h2); hi.w = *reinterpret_cast<const uint32_t*>(&h3);
    lo.x = *reinterpret_cast<const uint32_t*>(&l0); lo.y = *reinterpret_cast<const uint32_t*>(&l1);
    lo.z = *reinterpret_cast<const uint32_t*>(&l2); lo.w = *reinterpret_cast<const uint32_t*>(&l3);
    *reinterpret_cast<uint4*>(dst) = hi;
    *reinterpret_cast<uint4*>(dst + 64) = lo;
  }
  const __half2 gt = __hgt2(ovf, __floats2half2_rn(65504.f, 65504.f));
  const bool bad = __low2float(gt) != 0.f || __high2float(gt) != 0.f;
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0 && overflow) atomicAdd(overflow, 1u);
}

uint32_t* g_ss_timers = nullptr;  // FSFB_GEMM_TIMERS=1: role counters of the last launch (fsfb_debug_gemm_ss_timers)

// Launch helper called from fsfb_gather_gemm (gemm_tc.cu).  Returns 1 when the shape is not served by this kernel (the
// caller falls back to gemm_ts.cu / gemm_tc.cu), FSFB_OK after a launch, a negative status on errors.
// FSFB_GEMM_TIMERS=1: the zeroed [148][32] u32 counter block of this launch (nullptr otherwise)
int ss_timers_buffer(uint32_t** out, cudaStream_t st) {
  *out = nullptr;
  static const bool timed = [] { const char* e = getenv("FSFB_GEMM_TIMERS"); return e && atoi(e) != 0; }();
  if (timed) {
    if (!g_ss_timers) FSFB_CUDA(cudaMalloc(&g_ss_timers, (size_t)kNumSMs * 32 * 4));
    FSFB_CUDA(cudaMemsetAsync(g_ss_timers, 0, (size_t)kNumSMs * 32 * 4, st));
    *out = g_ss_timers;
  }
  return FSFB_OK;
}

// device address of the fp16-overflow counter on the current device (shared with gemm_lin.cu)
int ss_overflow_counter(unsigned int** out) {
  constexpr int kMaxDev = 64;
  static unsigned int* ptr[kMaxDev] = {nullptr};
  int dev = 0;
  FSFB_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= kMaxDev) {
    set_error("gather_gemm: device ordinal %d out of range", dev);
    return FSFB_ERR_BADARG;
  }
  if (!ptr[dev]) FSFB_CUDA(cudaGetSymbolAddress((void**)&ptr[dev], g_ss_overflow));
  *out = ptr[dev];
  return FSFB_OK;
}

int launch_gather_gemm_ss(TcParams& P, bool a_vec, bool a_split, float* workspace, size_t workspace_bytes, int splits,
                          const float* host_bias, const float* host_norm_w, const float* host_norm_b, cudaStream_t st) {
  const int n_pad = P.S.n_pad();
  if (!gemm_f16_enabled() || P.koff > 27) return 1;
  if (a_split && (P.cin % kGemmKChunk != 0 || ((uintptr_t)P.a & 15) != 0)) return 1;
  // column tiles: 128 wide; a single wide tile (single-buffered accumulators) for widths in (128, 256) that are not a
  // multiple of 128, so that the rows are gathered once and fused LayerNorms see the whole row
  int tile_w;
  if (n_pad <= 128) tile_w = n_pad;
  else if (n_pad % 128 == 0 && P.E.norm != FSFB_NORM_LAYERNORM) tile_w = 128;
  else if (n_pad <= 256) tile_w = n_pad;
  else return 1;
  P.ss_tile_w = tile_w;
  P.n_ct = (n_pad + tile_w - 1) / tile_w;
  P.splits = splits < 1 ? 1 : (splits > P.koff ? P.koff : splits);
  const int cpad = (n_pad + 127) & ~127;
  if (P.splits > 1) {
    if (tile_w > 128) return 1;
    const size_t need = (size_t)P.splits * (size_t)P.rows * cpad * sizeof(float);
    if (!workspace || workspace_bytes < need || cpad > 1024) {
      set_error("gather_gemm: split workspace too small (%zu given, %zu needed) or cout > 1024", workspace_bytes, need);
      return FSFB_ERR_CAPACITY;
    }
    P.partial = workspace;
  } else {
    P.partial = nullptr;
  }
  P.cpad = cpad;
  P.n_row_tiles = (int)ceil_div(P.rows, kTcRows);
  P.n_units = ceil_div(P.rows, kTcRows) * P.n_ct * P.splits;
  if (P.n_units >= (1ll << 31)) {
    set_error("gather_gemm: too many work units");
    return FSFB_ERR_BADARG;
  }
  int acc_cols = 32;
  while (acc_cols < tile_w) acc_cols <<= 1;
  P.ss_acc_cols = acc_cols;
  P.ss_bufs = 4 * acc_cols <= 512 ? 2 : 1;
  P.ss_stage_stride = ((tile_w + 31) & ~31) + 4;  // the block epilogue stages whole 32-column groups
  P.ss_w_slot = (uint32_t)align_up((size_t)tile_w * 128, 1024);
  const size_t nbr_bytes = P.nbr ? 2 * (size_t)P.koff * kTcRows * 4 : 0;
  const size_t staging = (size_t)kTcRows * P.ss_stage_stride * 4;
  const size_t vec = 3 * 256 * 4;
  const size_t budget = 227 * 1024;
  const int kc_n = P.S.kc();
  // ring depths: W first (a Linear layer keeps all its K chunks resident when they fit), then as many A slots as remain
  static const int w_cap = [] { const char* e = getenv("FSFB_SS_W_STAGES"); return e ? std::max(2, std::min(atoi(e), kSsMaxW)) : kSsMaxW; }();
  int w_stages = std::min(w_cap, std::max(2, std::min(kc_n * P.koff, kSsMaxW)));
  int a_stages = 0;
  for (;; --w_stages) {
    const size_t fixed = (size_t)w_stages * P.ss_w_slot + nbr_bytes + staging + vec + sizeof(SsShared) + 64;
    if (fixed < budget) a_stages = (int)std::min<size_t>(kSsMaxA, (budget - fixed) / kSsASlot);
    if (a_stages >= 3 || w_stages <= 2) break;
  }
  if (a_stages < 2) return 1;
  P.ss_a_stages = a_stages;
  P.ss_w_stages = w_stages;
  P.ss_off_w = (uint32_t)a_stages * kSsASlot;
  P.ss_off_nbr = P.ss_off_w + (uint32_t)w_stages * P.ss_w_slot;
  P.ss_off_stage = P.ss_off_nbr + (uint32_t)nbr_bytes;
  P.ss_off_vec = P.ss_off_stage + (uint32_t)staging;
  P.ss_off_sh = (uint32_t)align_up((size_t)P.ss_off_vec + vec, 16);
  const size_t smem = (size_t)P.ss_off_sh + sizeof(SsShared);
  if (smem > budget) return 1;

  constexpr int kMaxDev = 64;
  static unsigned launch_seq[kMaxDev] = {0};
  static unsigned sched_next[kMaxDev][kSsSchedSlots] = {{0}};
  static unsigned int* overflow_ptr[kMaxDev] = {nullptr};
  int dev = 0;
  FSFB_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= kMaxDev) {
    set_error("gather_gemm: device ordinal %d out of range", dev);
    return FSFB_ERR_BADARG;
  }
  if (!overflow_ptr[dev]) FSFB_CUDA(cudaGetSymbolAddress((void**)&overflow_ptr[dev], g_ss_overflow));
  P.ss_overflow = overflow_ptr[dev];
  P.sched_slot = (int)(launch_seq[dev]++ % kSsSchedSlots);
  P.sched_base = sched_next[dev][P.sched_slot];
  const unsigned grid = (unsigned)std::min<int64_t>(P.n_units, kNumSMs);
  sched_next[dev][P.sched_slot] += (unsigned)P.n_units + grid;
  {
    const int rc = ss_timers_buffer(&P.timers, st);
    if (rc != FSFB_OK) return rc;
  }

  static const bool hv_on = [] { const char* e = getenv("FSFB_GEMM_HV"); return !e || atoi(e) != 0; }();
  const bool hv = hv_on && P.n_ct == 1 && P.splits == 1 && tile_w <= 128 && (!P.E.bias || host_bias) && (!P.E.norm_w || host_norm_w) &&
                  (!P.E.norm_b || host_norm_b) && (P.E.act & 0xff) != FSFB_ACT_GELU;
  if (hv) {
    for (int c = 0; c < 128; ++c) {
      const bool in = c < P.S.cout;
      P.hv_bias[c] = (in && P.E.bias) ? host_bias[c] : 0.f;
      P.hv_w[c] = (in && P.E.norm_w) ? host_norm_w[c] : 1.f;
      P.hv_h[c] = (in && P.E.norm_b) ? host_norm_b[c] : 0.f;
    }
  }
  static bool attr = false;
  if (!attr) {
    FSFB_CUDA(cudaFuncSetAttribute((k_gather_gemm_ss<0, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget));
    FSFB_CUDA(cudaFuncSetAttribute((k_gather_gemm_ss<0, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget));
    FSFB_CUDA(cudaFuncSetAttribute((k_gather_gemm_ss<1, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget));
    FSFB_CUDA(cudaFuncSetAttribute((k_gather_gemm_ss<1, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget));
    FSFB_CUDA(cudaFuncSetAttribute((k_gather_gemm_ss<2, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget));
    FSFB_CUDA(cudaFuncSetAttribute((k_gather_gemm_ss<2, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget));
    FSFB_CUDA(cudaFuncSetAttribute((k_gather_gemm_ss<3, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget));
    FSFB_CUDA(cudaFuncSetAttribute((k_gather_gemm_ss<3, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget));
    FSFB_CUDA(cudaFuncSetAttribute((k_gather_gemm_ss<4, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget));
    FSFB_CUDA(cudaFuncSetAttribute((k_gather_gemm_ss<4, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget));
    attr = true;
  }
  // pre-split rows: 16-byte cp.async copies by four warps (LOAD 3).  Measured alternatives (profiles/r2_ncu_summary.md section 2):
  // register-staged 128-bit loads + plain stores by eight warps (LOAD 4, FSFB_GEMM_SPLIT_LDG=1) 389 vs 297 us on the 160 k-voxel
  // layer — the per-warp bookkeeping costs more issue slots than the cp.async form's lower copy rate (1,180 clk per 128-row stage
  // with nothing else running); TMA tile::gather4 2,510 clk per stage (tools/microbench/tma_gather4.cu)
  static const bool split_ldg = [] { const char* e = getenv("FSFB_GEMM_SPLIT_LDG"); return e && atoi(e) != 0; }();
  const int load = a_split ? (split_ldg ? 4 : 3) : (!a_vec ? 2 : (P.cin % kGemmKChunk == 0 ? 0 : 1));
  if (load == 4) {
    if (hv) FSFB_LAUNCH((k_gather_gemm_ss<4, true>), grid, kSsThreads, smem, st, P);
    else FSFB_LAUNCH((k_gather_gemm_ss<4, false>), grid, kSsThreads, smem, st, P);
  } else if (load == 3) {
    if (hv) FSFB_LAUNCH((k_gather_gemm_ss<3, true>), grid, kSsThreads, smem, st, P);
    else FSFB_LAUNCH((k_gather_gemm_ss<3, false>), grid, kSsThreads, smem, st, P);
  } else if (load == 0) {
    if (hv) FSFB_LAUNCH((k_gather_gemm_ss<0, true>), grid, kSsThreads, smem, st, P);
    else FSFB_LAUNCH((k_gather_gemm_ss<0, false>), grid, kSsThreads, smem, st, P);
  } else if (load == 1) {
    if (hv) FSFB_LAUNCH((k_gather_gemm_ss<1, true>), grid, kSsThreads, smem, st, P);
    else FSFB_LAUNCH((k_gather_gemm_ss<1, false>), grid, kSsThreads, smem, st, P);
  } else {
    if (hv) FSFB_LAUNCH((k_gather_gemm_ss<2, true>), grid, kSsThreads, smem, st, P);
    else FSFB_LAUNCH((k_gather_gemm_ss<2, false>), grid, kSsThreads, smem, st, P);
  }
  if (P.splits > 1) launch_splitk_epilogue(P, st);
  return FSFB_OK;
}

}  // namespace fsfb

extern "C" int fsfb_split_rows(const float* a, int64_t rows, int c, int64_t a_stride, void* out, void* stream) {
  using namespace fsfb;
  FSFB_CHECK_ARG(rows >= 0 && c >= 32 && c % 32 == 0 && a_stride >= c, "split_rows: c=%d must be a positive multiple of 32 (stride %lld)", c,
                 (long long)a_stride);
  if (rows == 0) return FSFB_OK;
  FSFB_CHECK_ARG(a && out, "split_rows: null pointer");
  FSFB_CHECK_ARG(((uintptr_t)a & 15) == 0 && a_stride % 4 == 0 && ((uintptr_t)out & 15) == 0, "split_rows: rows must be 16-byte aligned");
  unsigned int* ovf = nullptr;
  FSFB_CUDA(cudaGetSymbolAddress((void**)&ovf, g_ss_overflow));
  const int64_t total = rows * (c / 8);
  const int grid = (int)std::min<int64_t>(ceil_div(total, 256), (int64_t)kNumSMs * 16);
  FSFB_LAUNCH(k_split_rows, grid, 256, 0, (cudaStream_t)stream, a, rows, c, a_stride, (unsigned char*)out, ovf);
  return FSFB_OK;
}

// Diagnostics (tools/gemm_role_timers.py; not part of the product path): per-CTA role counters of the last launch made under
// FSFB_GEMM_TIMERS=1, out[148][32] u32.  Per CTA: [0..3] producer warp 0 {wait slot empty, convert + store, advance + issue
// gathers, stages}; [4..7] MMA {open unit + wait accumulators free, wait W, wait A, issue}; [8..12] epilogue warp 16 {open +
// residual / vectors, wait accumulators full, drain + epilogue math, rows out, units}; [31] producer total.
extern "C" int fsfb_debug_gemm_ss_timers(unsigned int* out) {
  using namespace fsfb;
  FSFB_CHECK_ARG(out != nullptr, "debug_gemm_ss_timers: null pointer");
  if (!g_ss_timers) {
    set_error("debug_gemm_ss_timers: no timed launch yet (set FSFB_GEMM_TIMERS=1)");
    return FSFB_ERR_BADARG;
  }
  FSFB_CUDA(cudaDeviceSynchronize());
  FSFB_CUDA(cudaMemcpy(out, g_ss_timers, (size_t)kNumSMs * 32 * 4, cudaMemcpyDeviceToHost));
  return FSFB_OK;
}

// Launches of the fp16-split gather-GEMM (this device) that met an input outside fp16 range since the library was loaded.
extern "C" int fsfb_gemm_f16_overflows(unsigned int* count) {
  using namespace fsfb;
  FSFB_CHECK_ARG(count != nullptr, "gemm_f16_overflows: null pointer");
  FSFB_CUDA(cudaDeviceSynchronize());
  FSFB_CUDA(cudaMemcpyFromSymbol(count, g_ss_overflow, sizeof(unsigned int)));
  return FSFB_OK;
}
