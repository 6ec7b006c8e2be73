// gemm_ts.cu — persistent gather-GEMM with the A operand in TENSOR MEMORY (tcgen05.mma "TS" form).
//
// Same contract as gemm_tc.cu (out[r] = epi(sum_k a[nbr[k][r]] @ w[k]^T)); serves every layer whose output
// width is <= 128 or a multiple of 128 (column tiles of 128), i.e. all sparse convolutions and almost every
// Linear of the stock FSF networks.
//
// One CTA per SM pulls work units from a global counter (heaviest row tiles first: fsfb_rulebook_row_order sorts
// rows by the number of present offsets, and units are handed out from the end of that order); a unit is (row tile
// of 128 output rows, column tile of <= 128 channels, offset split).  Measured on the previous one-tile-per-CTA kernel
// (round-1 experiments, profiles/r1_ncu_summary.md): launch + neighbour-tile prologue (~13 k clk) and the drain + epilogue
// (~23 k clk) of a tile cost more than its ~43 MMA stages (33 k clk) and nothing overlapped, because TMEM and
// shared memory admit one CTA per SM.  Here the roles run ahead of each other across units:
//   * warps 0-15  A producers (stage-striped groups of 4 warps, thread = row): gather 128 B of their row with
//     predicated 128-bit loads, split fp32 → tf32 hi/lo in registers, tcgen05.st into a 4-slot ring of TMEM columns.
//     The ring keeps running across unit boundaries: the gathers of unit i+1 fly while unit i drains.
//   * warps 20,21 MMA issuers (main accumulator a_hi*w_hi; correction accumulator a_lo*w_hi + a_hi*w_lo),
//     warp-uniform operands, one elected lane issues; tcgen05.commit recycles ring slots.
//   * warp 22     W loader: bulk-async copies of the pre-packed W blocks into a 4-slot shared-memory ring; a slot
//     that already holds the block it needs is not re-copied (Linear layers with K <= 128 load W once per CTA).
//   * warps 16-19 epilogue: fetch the NEXT unit id and prefetch its neighbour tile and active-offset mask into a
//     double-buffered table; then drain the accumulators of the current unit from TMEM (thread = lane = row) into a
//     shared-memory staging tile, hand the accumulators back to the MMA warps at once, and only then run the fused
//     bias / LayerNorm / affine / residual / activation epilogue with coalesced 512-byte row stores (each warp
//     finishes the 32 rows it staged itself, so a __syncwarp is the only synchronisation).
// TMEM columns: [0,acc) main accumulator, [acc,2acc) correction accumulator, then 4 x 64 columns of A slots
// (32 hi + 32 lo).  512 columns are allocated once per CTA.
//
// Split mode (P.splits > 1): the offsets are cut into P.splits contiguous ranges and each unit writes its raw
// partial sums to P.partial[split][row][cout_pad]; k_splitk_epilogue then sums the slabs in a fixed order
// (deterministic) and applies the epilogue.  Used when rows/128 x column tiles leaves most of the 148 SMs idle
// (the 1 k- and 11 k-row levels of the U-Net, the 1024-wide cluster heads).
#include <cstdlib>

#include <cuda_fp16.h>

#include "gemm_persist.cuh"

namespace fsfb {

constexpr int kTsProducers = 512;                    // warps 0-15 (warpgroups 0-3): warp w owns TMEM lane quarter w%4, ring slot w/4
constexpr int kTsEpiWarp = kTsProducers / 32;        // warps 16-19 (warpgroup 4): epilogue, TMEM lane quarter w%4
constexpr int kTsMmaWarp = kTsEpiWarp + 4;           // warps 20, 21 (warpgroup 5): MMA issuers
constexpr int kTsLoaderWarp = kTsMmaWarp + 2;        // warp 22: W loader
constexpr int kTsSchedWarp = kTsMmaWarp + 3;         // warp 23: pulls unit ids, prefetches their neighbour tiles
constexpr int kTsThreads = 768;
constexpr int kTsTableReaders = kTsProducers / 32 + 4 + 3;  // warps that read a unit's neighbour table / mask
constexpr int kTsAStages = 4;
constexpr uint32_t kTsWSlotBytes = 2 * 128 * 128;    // hi + lo of a 128-channel column tile, one K chunk
constexpr uint32_t kTsStageStride = 128 + 4;         // floats per staged accumulator row (bank-conflict-free 128-bit rows)

struct TsShared {
  uint64_t a_full[kTsAStages];
  uint64_t a_empty[kTsAStages];
  uint64_t nbr_full[2];
  uint64_t nbr_empty[2];
  uint64_t acc_full;
  uint64_t acc_empty;
  uint64_t res_full[4];     // residual rows of each epilogue warp have landed in its staging rows
  uint32_t tmem_base;
  uint32_t unit[2];         // unit id of the table buffer (>= n_units: no more work)
  uint32_t off_mask[2];     // its active offsets
  int32_t rows[2][kTcRows]; // output row of each tile row (-1: past the end)
};

// Work counters of the persistent launches: slot = launch sequence number mod kTsSchedSlots.  A counter is never reset:
// a launch draws exactly n_units + gridDim.x tickets (every CTA ends on one ticket past the end), so the host knows
// the value the next user of the slot starts from (P.sched_base) and no memset or clean-up kernel is needed.
constexpr int kTsSchedSlots = 256;
__device__ unsigned int g_ts_sched[kTsSchedSlots];


// Source of every "no neighbour" row: loads stay unconditional (no predicates, no zero-fill moves) and hit L1.
constexpr int kTsZeroRow = 2048;
__device__ __align__(32) float g_ts_zero_row[kTsZeroRow];


// AVEC: rows of `a` are 16-byte aligned; KFULL: additionally cin % 32 == 0 (every K chunk complete) → the gather is 8 plain
// 128-bit loads per thread and stage
// TIMED (FSFB_GEMM_TIMERS=1, tools/gemm_timers.py): per-role wait/work cycle counters, one row of 16 per CTA
#define TS_T0() uint32_t t0_ = TIMED ? (uint32_t)clock() : 0u
#define TS_ACC(var) do { if (TIMED) { const uint32_t t1_ = (uint32_t)clock(); var += t1_ - t0_; t0_ = t1_; } } while (0)
// HV: the per-channel epilogue vectors travel in the kernel parameters (cout <= 128, host copies given): the epilogue reads
// them as constant-bank operands of its FMAs, i.e. with no load instruction at all
// F16 (FSFB_GEMM_F16=1, needs KFULL): fp16-split operands — a = hi + lo / 2048 with hi = fp16(a), lo = fp16((a - hi) * 2048),
// the same 22 mantissa bits as the tf32 split — through kind::f16 MMAs (K = 16 per instruction, so 6 instead of 12 per stage
// and half the TMEM / shared-memory operand bytes); the correction accumulator is scaled back by 2^-11 in the epilogue.
template <bool AVEC, bool KFULL, bool TIMED = false, bool HV = false, bool F16 = false>
__global__ void __launch_bounds__(kTsThreads, 1) k_gather_gemm_ts(const __grid_constant__ TcParams P) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t base = smem_u32(smem_raw);
  const uint32_t nbr_bytes = (uint32_t)P.koff * kTcRows * 4u;
  const uint32_t s_nbr = base + kTsAStages * kTsWSlotBytes;        // [2][koff][128] i32
  const uint32_t s_stage = s_nbr + 2 * nbr_bytes;                  // [128][kTsStageStride] f32: drained accumulators, finished in place
  const uint32_t s_vec = s_stage + kTcRows * kTsStageStride * 4u;  // [3][128] f32: bias, norm_w, norm_b of the column tile
  TsShared* sh = reinterpret_cast<TsShared*>(smem_raw + (size_t)kTsAStages * kTsWSlotBytes + 2 * (size_t)nbr_bytes +
                                             (size_t)kTcRows * kTsStageStride * 4 + 3 * 128 * 4);

  if (tid == 0) {
    if (base & 1023u) __trap();
    for (int s = 0; s < kTsAStages; ++s) {
      mbar_init(smem_u32(&sh->a_full[s]), 4 + 1);  // the 4 warps of the slot's producer group + the W loader (with its tx bytes)
      mbar_init(smem_u32(&sh->a_empty[s]), 2);     // both MMA issuers commit
    }
    mbar_init(smem_u32(&sh->nbr_full[0]), 1);      // the scheduler warp
    mbar_init(smem_u32(&sh->nbr_full[1]), 1);
    mbar_init(smem_u32(&sh->nbr_empty[0]), kTsTableReaders);  // every reader of the table (the epilogue warps read the mask)
    mbar_init(smem_u32(&sh->nbr_empty[1]), kTsTableReaders);
    mbar_init(smem_u32(&sh->acc_full), 2);
    mbar_init(smem_u32(&sh->acc_empty), 4);
    for (int w = 0; w < 4; ++w) mbar_init(smem_u32(&sh->res_full[w]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  uint32_t acc_cols = 32;
  {
    const int n_max = min(128, P.S.n_pad());
    while ((int)acc_cols < n_max) acc_cols <<= 1;
  }
  constexpr uint32_t tmem_cols = 512;
  if (warp == kTsMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sh->tmem_base)),
                 "r"(tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = sh->tmem_base;
  const uint32_t tmem_a = tmem_d + 2 * acc_cols;  // A ring: slot s at +64*s (hi 32 cols | lo 32 cols)
  const int kc_n = P.S.kc();

  // every role opens its iter-th unit from the table the epilogue warps filled: unit id + active-offset mask
  auto open_unit = [&](int iter, TsUnit& U, uint32_t& mask) -> bool {
    const int b = iter & 1;
    mbar_wait(smem_u32(&sh->nbr_full[b]), (uint32_t)(iter >> 1) & 1u);
    if (!ts_unit(P, lds_u32(smem_u32(&sh->unit[b])), U)) return false;
    mask = lds_u32(smem_u32(&sh->off_mask[b])) & U.k_keep;
    return true;
  };

  // Register budget: 768 threads x 80 at launch; the MMA/loader warpgroup hands registers back and the producer
  // warpgroups take them (setmaxnreg is per warpgroup), so the gather prefetch (32 registers) is not spilled.
  if (tid < kTsProducers) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 88;" ::: "memory");
    // ================= A producers =================
    // group g = warps 4g..4g+3 (one per TMEM lane quarter) produces the stages whose global index is g mod 4 into ring
    // slot g: a thread owns one ROW of the tile and gathers its whole 128-byte K chunk (8 128-bit loads).
    const int grp = warp >> 2, q = warp & 3;
    const int row = 32 * q + lane;
    const uint32_t ta = tmem_a + ((uint32_t)(32 * q) << 16) + (uint32_t)(64 * grp);
    const uint32_t full_bar = smem_u32(&sh->a_full[grp]), empty_bar = smem_u32(&sh->a_empty[grp]);
    // flat walk over (unit, stage) restricted to this group's stages
    int iter = -1, j = 0, n_act = 0;  // open unit (none yet), this group's next stage in it, its stage count
    uint32_t gbase = 0;               // global stage count before the open unit (mod 4 is what matters)
    TsCursor c;
    c.rem = 0;
    c.kc = 0;
    uint32_t nbr_b = 0;
    auto next_stage = [&]() -> bool {
      while (j >= n_act) {  // open the next unit
        if (iter >= 0) {
          __syncwarp();  // this warp has read everything it needs from the unit's neighbour table
          if (lane == 0) mbar_arrive(smem_u32(&sh->nbr_empty[iter & 1]));
          gbase += (uint32_t)n_act;
        }
        ++iter;
        TsUnit U;
        uint32_t m;
        if (!open_unit(iter, U, m)) return false;
        n_act = __popc(m) * kc_n;
        j = (int)((grp - gbase) & 3u);
        c.init(m, j, kc_n);
        nbr_b = s_nbr + (uint32_t)(iter & 1) * nbr_bytes;
      }
      return true;
    };
    auto load_stage = [&](float4(&v)[8]) {
      int32_t src = lds_i32(nbr_b + (uint32_t)(c.k() * kTcRows + row) * 4u);
      if (src >= P.a_rows || (P.debug & 1)) src = -1;
      const int col0 = c.kc * kGemmKChunk;
      if (KFULL) {  // 32-byte aligned rows: four 256-bit loads (half the L1 requests of 128-bit ones: the gather is request-bound)
        const float* g = (src >= 0 ? P.a + (int64_t)src * P.a_stride : g_ts_zero_row) + col0;
#pragma unroll
        for (int jj = 0; jj < 4; ++jj)
          asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                       : "=f"(v[2 * jj].x), "=f"(v[2 * jj].y), "=f"(v[2 * jj].z), "=f"(v[2 * jj].w), "=f"(v[2 * jj + 1].x),
                         "=f"(v[2 * jj + 1].y), "=f"(v[2 * jj + 1].z), "=f"(v[2 * jj + 1].w)
                       : "l"(g + 8 * jj));
      } else {
        const float* g0 = P.a + (int64_t)(src >= 0 ? src : 0) * P.a_stride + col0;
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          const int col = col0 + 4 * jj;
          const bool ok = src >= 0 && col < P.cin;
          const float* g = ok ? g0 + 4 * jj : P.a;
          if (AVEC && col + 4 <= P.cin) {
            v[jj] = ldg_pred_f4(g, ok);
          } else {
            v[jj].x = ldg_pred_f1(g, ok);
            v[jj].y = ldg_pred_f1(g + 1, ok && col + 1 < P.cin);
            v[jj].z = ldg_pred_f1(g + 2, ok && col + 2 < P.cin);
            v[jj].w = ldg_pred_f1(g + 3, ok && col + 3 < P.cin);
          }
        }
      }
    };
    float4 cur[8];
    uint32_t tm_empty = 0, tm_conv = 0, tm_next = 0, tm_n = 0;
    const uint32_t tm_start = TIMED ? (uint32_t)clock() : 0u;
    bool have = next_stage();
    if (have) load_stage(cur);
    uint32_t ph = 0;
    while (have) {
      TS_T0();
      if (lane == 0) mbar_wait(empty_bar, ph ^ 1u);
      __syncwarp();
      tc_fence_after();
      TS_ACC(tm_empty);
      // hi = x with the 13 low mantissa bits cleared (what the tensor core keeps of a tf32 operand),
      // lo = x - hi exactly (the MMA truncates it to tf32 itself): 2 ALU ops per element
      // (two halves of 16 columns each keep the transient registers at 16)
      if constexpr (F16) {
        // column c of the hi half holds the chunk's inputs (2c, 2c+1) as fp16 (low half = even input); the lo half holds
        // (x - hi) * 2048, exact in fp32 before its own rounding to fp16
        uint32_t t16[16];
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          const __half2 h0 = __floats2half2_rn(cur[jj].x, cur[jj].y), h1 = __floats2half2_rn(cur[jj].z, cur[jj].w);
          t16[2 * jj + 0] = *reinterpret_cast<const uint32_t*>(&h0);
          t16[2 * jj + 1] = *reinterpret_cast<const uint32_t*>(&h1);
        }
        tc_st16(ta, t16);
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          const float2 b0 = __half22float2(*reinterpret_cast<const __half2*>(&t16[2 * jj + 0]));
          const float2 b1 = __half22float2(*reinterpret_cast<const __half2*>(&t16[2 * jj + 1]));
          const __half2 l0 = __floats2half2_rn((cur[jj].x - b0.x) * kF16LoScale, (cur[jj].y - b0.y) * kF16LoScale);
          const __half2 l1 = __floats2half2_rn((cur[jj].z - b1.x) * kF16LoScale, (cur[jj].w - b1.y) * kF16LoScale);
          t16[2 * jj + 0] = *reinterpret_cast<const uint32_t*>(&l0);
          t16[2 * jj + 1] = *reinterpret_cast<const uint32_t*>(&l1);
        }
        tc_st16(ta + 16, t16);
      } else {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t t16[16];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          t16[4 * jj + 0] = __float_as_uint(cur[4 * h + jj].x) & 0xFFFFE000u;
          t16[4 * jj + 1] = __float_as_uint(cur[4 * h + jj].y) & 0xFFFFE000u;
          t16[4 * jj + 2] = __float_as_uint(cur[4 * h + jj].z) & 0xFFFFE000u;
          t16[4 * jj + 3] = __float_as_uint(cur[4 * h + jj].w) & 0xFFFFE000u;
        }
        tc_st16(ta + 16 * h, t16);
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          t16[4 * jj + 0] = __float_as_uint(cur[4 * h + jj].x - __uint_as_float(t16[4 * jj + 0]));
          t16[4 * jj + 1] = __float_as_uint(cur[4 * h + jj].y - __uint_as_float(t16[4 * jj + 1]));
          t16[4 * jj + 2] = __float_as_uint(cur[4 * h + jj].z - __uint_as_float(t16[4 * jj + 2]));
          t16[4 * jj + 3] = __float_as_uint(cur[4 * h + jj].w - __uint_as_float(t16[4 * jj + 3]));
        }
        tc_st16(ta + 32 + 16 * h, t16);
      }
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(full_bar);
      TS_ACC(tm_conv);
      // the next stage of this group (possibly in the next unit) is four stage-times away: its gathers fly meanwhile
      j += kTsAStages;
      c.step(kTsAStages, kc_n);
      have = next_stage();
      if (have) load_stage(cur);
      ph ^= 1u;
      TS_ACC(tm_next);
      ++tm_n;
    }
    if (TIMED && P.timers && (tid & 127) == 0) {
      uint32_t* t = P.timers + (size_t)blockIdx.x * 32 + 4 * grp;
      t[0] = tm_empty; t[1] = tm_conv; t[2] = tm_next; t[3] = tm_n;
      if (grp == 0) P.timers[(size_t)blockIdx.x * 32 + 31] = (uint32_t)clock() - tm_start;
    }
  } else if (warp >= kTsMmaWarp) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;" ::: "memory");
    if (warp == kTsMmaWarp || warp == kTsMmaWarp + 1) {
      // ================= two MMA issuer warps =================
      // first warp: main accumulator (a_hi * w_hi); second warp: correction accumulator (a_lo * w_hi + a_hi * w_lo).
      // The whole warp runs the loop with warp-uniform operands (laundered through __shfl_sync so ptxas keeps
      // descriptors / TMEM addresses in uniform registers) and one elected lane issues.
      const bool is_main = warp == kTsMmaWarp;
      const uint32_t u_tmem_d = __shfl_sync(0xffffffffu, tmem_d, 0);
      const uint32_t u_tmem_a = __shfl_sync(0xffffffffu, tmem_a, 0);
      const uint32_t u_base = __shfl_sync(0xffffffffu, base, 0);
      const uint32_t d_acc = is_main ? u_tmem_d : u_tmem_d + acc_cols;
      const uint32_t bar_full0 = __shfl_sync(0xffffffffu, smem_u32(&sh->a_full[0]), 0);
      const uint32_t bar_empty0 = __shfl_sync(0xffffffffu, smem_u32(&sh->a_empty[0]), 0);
      int s = 0;
      uint32_t ph = 0;
      TsUnit U;
      uint32_t m0;
      uint32_t tm_open = 0, tm_acc = 0, tm_full = 0, tm_issue = 0;
      TS_T0();
      for (int iter = 0; open_unit(iter, U, m0); ++iter) {
        TS_ACC(tm_open);
        const uint32_t m = __shfl_sync(0xffffffffu, m0, 0);
        const int n_sub = __shfl_sync(0xffffffffu, U.n_sub, 0);
        if (lane == 0) mbar_arrive(smem_u32(&sh->nbr_empty[iter & 1]));
        const int n_active = __popc(m) * kc_n;
        const uint32_t idesc = F16 ? make_idesc_f16(n_sub) : make_idesc_tf32(n_sub);
        if (iter > 0) {  // the epilogue has read the previous unit's accumulators out of TMEM
          mbar_wait(smem_u32(&sh->acc_empty), (uint32_t)(iter - 1) & 1u);
          tc_fence_after();
        }
        TS_ACC(tm_acc);
        StageCursor c;
        c.init(m);
        for (int it = 0; it < n_active; ++it) {
          if (lane == 0) mbar_wait(bar_full0 + 8u * s, ph);
          __syncwarp();
          tc_fence_after();
          TS_ACC(tm_full);
          const uint32_t w_hi = u_base + (uint32_t)s * kTsWSlotBytes;
          const uint32_t w_lo = w_hi + (uint32_t)n_sub * 128u;
          const uint32_t a_hi = u_tmem_a + (uint32_t)(64 * s), a_lo = a_hi + 32;
          const int k_valid = min(kGemmKChunk, P.cin - c.kc * kGemmKChunk);
          const int ksteps = (k_valid + 7) >> 3;
          uint32_t elected;
          asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(elected));
          if (elected) {
            if constexpr (F16) {
              // slot layout: A hi in columns [0,16), lo in [16,32); W rows = [hi 64 bytes | lo 64 bytes]; two K = 16 steps
              if (!(P.debug & 4)) {
                const uint32_t a_lo16 = a_hi + 16;
                if (is_main) {
#pragma unroll
                  for (int kk = 0; kk < 2; ++kk)
                    tc_mma_f16_ts(d_acc, a_hi + 8 * kk, make_sw128_desc(w_hi + 32u * kk), idesc, (it > 0 || kk > 0) ? 1u : 0u);
                } else {
#pragma unroll
                  for (int kk = 0; kk < 2; ++kk) {
                    tc_mma_f16_ts(d_acc, a_lo16 + 8 * kk, make_sw128_desc(w_hi + 32u * kk), idesc, (it > 0 || kk > 0) ? 1u : 0u);
                    tc_mma_f16_ts(d_acc, a_hi + 8 * kk, make_sw128_desc(w_hi + 64u + 32u * kk), idesc, 1u);
                  }
                }
              }
            } else
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
            if (it == n_active - 1) tc_commit(smem_u32(&sh->acc_full));
          }
          __syncwarp();
          c.next(kc_n);
          if (++s == kTsAStages) {
            s = 0;
            ph ^= 1u;
          }
          TS_ACC(tm_issue);
        }
        if (n_active == 0 && lane == 0) mbar_arrive(smem_u32(&sh->acc_full));  // nothing to wait for: the epilogue writes zeros
      }
      if (TIMED && P.timers && is_main && lane == 0) {
        uint32_t* t = P.timers + (size_t)blockIdx.x * 32 + 16;
        t[0] = tm_open; t[1] = tm_acc; t[2] = tm_full; t[3] = tm_issue;
      }
    } else if (warp == kTsLoaderWarp && lane == 0) {
      // ================= W loader: bulk-async copies into the slot ring =================
      int s = 0;
      uint32_t ph = 0;
      int tag[kTsAStages] = {-1, -1, -1, -1};
      TsUnit U;
      uint32_t m;
      for (int iter = 0; open_unit(iter, U, m); ++iter) {
        mbar_arrive(smem_u32(&sh->nbr_empty[iter & 1]));
        const int n_active = __popc(m) * kc_n;
        const int nt256 = U.ct >> 1;
        const uint32_t n_w256 = (uint32_t)P.S.n_w(nt256);
        const uint32_t sub_off = (uint32_t)(U.ct & 1) * 128u * 128u;
        const uint32_t sub_bytes = (uint32_t)U.n_sub * 128u;
        StageCursor c;
        c.init(m);
        for (int it = 0; it < n_active; ++it) {
          mbar_wait(smem_u32(&sh->a_empty[s]), ph ^ 1u);
          const int want = (U.ct * P.koff + c.k) * kc_n + c.kc;
          if (F16 && tag[s] != want && !(P.debug & 2)) {  // one 128-byte row per output channel holds hi and lo
            const unsigned char* blk = P.w_packed + P.S.f16_block_offset(nt256, c.k, c.kc) + sub_off;
            mbar_expect_tx(smem_u32(&sh->a_full[s]), sub_bytes);
            bulk_g2s(base + (uint32_t)s * kTsWSlotBytes, blk, sub_bytes, smem_u32(&sh->a_full[s]));
            tag[s] = want;
          }
          if (tag[s] != want && !(P.debug & 2)) {
            const unsigned char* blk = P.w_packed + P.S.block_offset(nt256, c.k, c.kc) + sub_off;
            const uint32_t dst = base + (uint32_t)s * kTsWSlotBytes;
            mbar_expect_tx(smem_u32(&sh->a_full[s]), 2 * sub_bytes);
            bulk_g2s(dst, blk, sub_bytes, smem_u32(&sh->a_full[s]));                                      // hi rows
            bulk_g2s(dst + sub_bytes, blk + (size_t)n_w256 * 128u, sub_bytes, smem_u32(&sh->a_full[s]));  // lo rows
            tag[s] = want;
          }
          mbar_arrive(smem_u32(&sh->a_full[s]));
          c.next(kc_n);
          if (++s == kTsAStages) {
            s = 0;
            ph ^= 1u;
          }
        }
      }
    } else if (warp == kTsSchedWarp) {
      // ================= scheduler: next unit id + its neighbour tile and active-offset mask → table buffer =================
      for (int iter = 0;; ++iter) {
        const int b = iter & 1;
        const uint32_t nb = s_nbr + (uint32_t)b * nbr_bytes;
        if (iter >= 2) mbar_wait(smem_u32(&sh->nbr_empty[b]), (uint32_t)((iter >> 1) + 1) & 1u);  // readers of unit iter-2 are done
        uint32_t u = 0;
        if (lane == 0) u = atomicAdd(&g_ts_sched[P.sched_slot], 1u) - P.sched_base;
        u = __shfl_sync(0xffffffffu, u, 0);
        TsUnit U;
        const bool have = ts_unit(P, u, U);
        uint32_t my_mask = 0;
        if (have) {
          // table entries go global → shared with 4-byte cp.async copies (no registers, all 4 x koff of a lane in flight
          // at once); rows past the end read a -1 from the zero-size form.  Producers range-check the entries themselves.
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
            sh->rows[b][r_l] = (int32_t)r[rq];
            if (P.nbr && r[rq] >= 0) {
              for (int k = 0; k < P.koff; ++k)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(nb + (uint32_t)(k * kTcRows + r_l) * 4u),
                             "l"(P.nbr + (int64_t)k * P.rows + r[rq])
                             : "memory");
            } else {
              for (int k = 0; k < P.koff; ++k) sts_i32(nb + (uint32_t)(k * kTcRows + r_l) * 4u, P.nbr ? -1 : (int32_t)r[rq]);
            }
          }
          asm volatile("cp.async.wait_all;" ::: "memory");
          __syncwarp();
          for (int k = 0; k < P.koff; ++k) {
            bool any = false;
#pragma unroll
            for (int rq = 0; rq < 4; ++rq) {
              const int32_t src = lds_i32(nb + (uint32_t)(k * kTcRows + 32 * rq + lane) * 4u);
              any |= src >= 0 && src < P.a_rows;
            }
            if (any) my_mask |= 1u << k;
          }
          my_mask = __reduce_or_sync(0xffffffffu, my_mask);
        }
        if (lane == 0) {
          sh->unit[b] = u;
          sh->off_mask[b] = my_mask;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&sh->nbr_full[b]));
        if (!have) break;
      }
    }
  } else {
    // ================= epilogue warps =================
    // Measured (tools/gemm_timers.py): while the producers' gathers keep the load/store unit busy, every memory
    // instruction of these warps (shared or global) queues ~100 clk behind them, and an in-order dependent chain exposes
    // that delay each time.  Hence, in fast mode (16-byte aligned output rows, channel count a multiple of 4):
    //   * thread = TMEM lane = row applies the whole epilogue in registers, 8 columns at a time; the per-channel vectors
    //     of those columns are requested from shared memory BEFORE the tcgen05.ld wait, so their queue delay overlaps it;
    //   * the finished row is left in the staging tile and leaves with ONE bulk-async copy (cp.async.bulk shared → global)
    //     issued by the row's own lane; residual rows arrive the same way while the unit's MMAs still run.
    // No global load/store instruction is issued here.  Slow mode (odd widths / unaligned rows): raw sums are staged and
    // finished lanes-over-channels by warp_row_epilogue with ordinary stores.
    const int q = warp & 3;
    const int r_l = 32 * q + lane;
    const uint32_t t_row = tmem_d + ((uint32_t)(32 * q) << 16);
    const uint32_t my_row = s_stage + (uint32_t)r_l * kTsStageStride * 4u;
    const uint32_t my_stage = s_stage + (uint32_t)(32 * q) * kTsStageStride * 4u;  // the 32 rows this warp drains and finishes
    const uint32_t res_bar = smem_u32(&sh->res_full[q]);
    const Epilogue& E = P.E;
    const int act = E.act & 0xff;
    const bool post = (E.act & FSFB_RESIDUAL_POST) != 0;
    const bool res_vec = !E.residual || (((uintptr_t)E.residual % 16 == 0) && (E.residual_stride % 4 == 0));
    uint32_t tm_pre = 0, tm_accf = 0, tm_p1 = 0, tm_p2 = 0, tm_units = 0;
    uint32_t res_ph = 0;
    int vec_ct = -1;  // column tile whose per-channel vectors are staged in s_vec
    TsUnit U;
    uint32_t m;
    for (int iter = 0; open_unit(iter, U, m); ++iter) {
      TS_T0();
      const int64_t r_cur = sh->rows[iter & 1][r_l] < 0 ? P.rows : (int64_t)sh->rows[iter & 1][r_l];
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&sh->nbr_empty[iter & 1]));
      TS_ACC(tm_pre);
      const bool have_acc = m != 0;
      const int c0 = U.ct * 128;
      const int c_n = min(U.n_sub, P.S.cout - c0);  // real channels in this column tile
      const bool split = P.splits > 1;
      const bool fast = split || ((c_n & 3) == 0 && P.out_vec && res_vec);
      const bool fused = fast && !split;
      const bool valid = r_cur < P.rows && !(P.debug & 32);
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
        const int c = c0 + r_l;
        const bool in = r_l < c_n;
        float* sv = reinterpret_cast<float*>(smem_raw + (s_vec - base));
        sv[r_l] = (in && E.bias) ? __ldg(E.bias + c) : 0.f;
        sv[128 + r_l] = (in && E.norm_w) ? __ldg(E.norm_w + c) : 1.f;
        sv[256 + r_l] = (in && E.norm_b) ? __ldg(E.norm_b + c) : 0.f;
        asm volatile("bar.sync 1, 128;" ::: "memory");
        vec_ct = U.ct;
      }
      mbar_wait(smem_u32(&sh->acc_full), (uint32_t)iter & 1u);
      tc_fence_after();
      TS_ACC(tm_accf);
      const bool blk32 = split || (HV && fused);
      if (blk32) {
        if (use_res) {
          mbar_wait(res_bar, res_ph);
          res_ph ^= 1u;
        }
        const uint32_t ae = smem_u32(&sh->acc_empty);
        const int nrm = (HV && fused) ? E.norm : FSFB_NORM_NONE;
        const int ac = (HV && fused) ? act : FSFB_ACT_NONE;
        const int cn = (HV && fused) ? c_n : 0;  // raw sums (offset splits): no column is finished here
#define TS_EPI(N, A, PO, UN) ts_epi32<N, A, PO, UN, F16>(P, t_row, acc_cols, my_row, cn, U.n_sub, have_acc, use_res, valid, ae, lane)
        if (ac == FSFB_ACT_RELU) {
          if (nrm == FSFB_NORM_LAYERNORM) { if (post) TS_EPI(FSFB_NORM_LAYERNORM, FSFB_ACT_RELU, true, true); else TS_EPI(FSFB_NORM_LAYERNORM, FSFB_ACT_RELU, false, true); }
          else if (nrm == FSFB_NORM_AFFINE) { if (post) TS_EPI(FSFB_NORM_AFFINE, FSFB_ACT_RELU, true, true); else TS_EPI(FSFB_NORM_AFFINE, FSFB_ACT_RELU, false, true); }
          else { if (post) TS_EPI(FSFB_NORM_NONE, FSFB_ACT_RELU, true, true); else TS_EPI(FSFB_NORM_NONE, FSFB_ACT_RELU, false, true); }
        } else {
          if (nrm == FSFB_NORM_LAYERNORM) TS_EPI(FSFB_NORM_LAYERNORM, FSFB_ACT_NONE, false, true);
          else if (nrm == FSFB_NORM_AFFINE) TS_EPI(FSFB_NORM_AFFINE, FSFB_ACT_NONE, false, true);
          else TS_EPI(FSFB_NORM_NONE, FSFB_ACT_NONE, false, true);
        }
#undef TS_EPI
      } else {
      float v[8];
      auto ld_issue = [&](int cb, uint32_t (&a)[8], uint32_t (&b)[8]) {  // warp-collective; pair with ld_wait
        if (have_acc) {
          tc_ld8_nowait(t_row + cb, a);
          tc_ld8_nowait(t_row + acc_cols + cb, b);
        }
      };
      auto ld_wait = [&](const uint32_t (&a)[8], const uint32_t (&b)[8]) {
        if (have_acc) {
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int jj = 0; jj < 8; ++jj)
            v[jj] = F16 ? fmaf(__uint_as_float(b[jj]), 1.f / kF16LoScale, __uint_as_float(a[jj])) : __uint_as_float(a[jj]) + __uint_as_float(b[jj]);
        } else {
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) v[jj] = 0.f;
        }
      };
      const bool has_bias = E.bias != nullptr;
      float mean = 0.f, rstd = 1.f;
      if (fused && E.norm == FSFB_NORM_LAYERNORM) {
        // LayerNorm statistics of this thread's row (two-pass, the whole row is in this tile: cout <= 128 enforced on the
        // host); build_mlp's normed layers carry no bias, so the bias reads are the rare path
        float sum = 0.f;
        for (int cb = 0; cb < c_n; cb += 8) {
          uint32_t a[8], b[8];
          ld_issue(cb, a, b);
          ld_wait(a, b);
#pragma unroll
          for (int jj = 0; jj < 8; jj += 4) {
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
            if (cb + jj < c_n) {
              const float d0 = v[jj] + bb.x - mean, d1 = v[jj + 1] + bb.y - mean, d2 = v[jj + 2] + bb.z - mean, d3 = v[jj + 3] + bb.w - mean;
              qq += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
            }
          }
        }
        rstd = 1.f / sqrtf(qq / (float)c_n + E.eps);
      }
      if (use_res) {
        mbar_wait(res_bar, res_ph);
        res_ph ^= 1u;
      }
      // ---- thread = TMEM lane = row: accumulators (+ fused epilogue in fast mode) → staging row, 8 columns at a time ----
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
              vw4[t] = lds_f4(s_vec + (uint32_t)(128 + cb + 4 * t) * 4u);
              vh4[t] = lds_f4(s_vec + (uint32_t)(256 + cb + 4 * t) * 4u);
            }
            if (use_res && valid && cb + 4 * t < c_n) g4[t] = lds_f4(srow + 16 * t);
          }
        }
        ld_wait(a, b);
        if (cb + 8 >= U.n_sub) {  // last TMEM read of this unit: the MMA warps may start the next unit
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&sh->acc_empty));
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
      TS_ACC(tm_p1);
      // ---- rows leave shared memory ----
      if (fast) {
        fence_proxy_async();  // this thread's staging writes (generic proxy) before its bulk copy's reads (async proxy)
        if (valid) {
          float* dst = split ? P.partial + ((int64_t)U.sp * P.rows + r_cur) * P.cpad + c0 : P.out + r_cur * P.out_stride + c0;
          const uint32_t bytes = (uint32_t)(split ? U.n_sub : c_n) * 4u;
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(my_row), "r"(bytes) : "memory");
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      } else if (!(P.debug & 32)) {
        __syncwarp();  // this warp's staging rows are complete
        Epilogue Es = E;
        if (Es.bias) Es.bias += c0;
        if (Es.norm_w) Es.norm_w += c0;
        if (Es.norm_b) Es.norm_b += c0;
        if (Es.residual) Es.residual += c0;
        for (int i = 0; i < 32; ++i) {
          const int64_t r = __shfl_sync(0xffffffffu, r_cur, i);
          if (r >= P.rows) continue;  // warp-uniform
          const float* xs = reinterpret_cast<const float*>(smem_raw + (my_stage - base) + (size_t)i * kTsStageStride * 4);
          warp_row_epilogue(xs, c_n, Es, r, P.out + r * P.out_stride + c0);
        }
        __syncwarp();  // reads of the staging rows finish before the next unit overwrites them
      }
      TS_ACC(tm_p2);
      ++tm_units;
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // the last rows are on their way out before the CTA exits
    if (TIMED && P.timers && warp == kTsEpiWarp && lane == 0) {
      uint32_t* t = P.timers + (size_t)blockIdx.x * 32 + 20;
      t[0] = tm_pre; t[1] = tm_accf; t[2] = tm_p1; t[3] = tm_p2; t[4] = tm_units;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kTsMmaWarp) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(tmem_cols) : "memory");
  }
}

// Sum of the split slabs + the fused epilogue (one warp per output row, fixed summation order).
__global__ void __launch_bounds__(256) k_splitk_epilogue(const float* __restrict__ partial, int splits, int64_t rows, int cpad,
                                                         int cout, Epilogue E, float* __restrict__ out, int64_t out_stride) {
  __shared__ float s_x[8][1024 + 32];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int64_t r = (int64_t)blockIdx.x * 8 + w; r < rows; r += (int64_t)gridDim.x * 8) {
    for (int c = lane * 4; c < cpad; c += 128) {
      float4 s = *reinterpret_cast<const float4*>(partial + r * cpad + c);
      for (int sp = 1; sp < splits; ++sp) {
        const float4 t = *reinterpret_cast<const float4*>(partial + ((int64_t)sp * rows + r) * cpad + c);
        s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
      }
      *reinterpret_cast<float4*>(&s_x[w][c]) = s;
    }
    __syncwarp();
    warp_row_epilogue(s_x[w], cout, E, r, out + r * out_stride);
    __syncwarp();
  }
}

int launch_splitk_epilogue(const TcParams& P, cudaStream_t st) {
  const int g2 = (int)std::min<int64_t>(ceil_div(P.rows, 8), (int64_t)kNumSMs * 8);
  FSFB_LAUNCH(k_splitk_epilogue, g2, 256, 0, st, P.partial, P.splits, P.rows, P.cpad, P.S.cout, P.E, P.out, P.out_stride);
  return FSFB_OK;
}

uint32_t* g_ts_timers = nullptr;  // FSFB_GEMM_TIMERS=1: counters of the last timed launch (fsfb_debug_gemm_timers)

// Launch helper called from fsfb_gather_gemm (gemm_tc.cu).
int launch_gather_gemm_ts(TcParams& P, bool a_vec, float* workspace, size_t workspace_bytes, int splits, const float* host_bias,
                          const float* host_norm_w, const float* host_norm_b, cudaStream_t st) {
  const int n_pad = P.S.n_pad();
  const int cpad = (n_pad + 127) & ~127;
  P.n_ct = (n_pad + 127) / 128;
  P.splits = splits < 1 ? 1 : (splits > P.koff ? P.koff : splits);
  if (P.splits > 1) {
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
  // the counters are __device__ variables, i.e. one set per GPU: the host-side ticket bookkeeping is kept per device too
  // (a process that launches on cuda:1 after cuda:0 must not carry cuda:0's totals over)
  constexpr int kMaxDev = 64;
  static unsigned launch_seq[kMaxDev] = {0};
  static unsigned sched_next[kMaxDev][kTsSchedSlots] = {{0}};  // tickets drawn so far from each slot (wraps with the device counter)
  int dev = 0;
  FSFB_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= kMaxDev) {
    set_error("gather_gemm: device ordinal %d out of range", dev);
    return FSFB_ERR_BADARG;
  }
  P.sched_slot = (int)(launch_seq[dev]++ % kTsSchedSlots);
  P.sched_base = sched_next[dev][P.sched_slot];
  P.n_units = ceil_div(P.rows, kTcRows) * P.n_ct * P.splits;
  if (P.n_units >= (1ll << 31)) {
    set_error("gather_gemm: too many work units");
    return FSFB_ERR_BADARG;
  }
  const size_t smem = (size_t)kTsAStages * kTsWSlotBytes + 2 * (size_t)P.koff * kTcRows * 4 + (size_t)kTcRows * kTsStageStride * 4 + 3 * 128 * 4 + sizeof(TsShared);
  const size_t budget = 227 * 1024;
  if (smem > budget) {
    set_error("gather_gemm: tile does not fit shared memory");
    return FSFB_ERR_BADARG;
  }
  // host copies of every per-channel vector that is present, one column tile, no offset split: vectors ride in the parameters
  static const bool hv_on = [] { const char* e = getenv("FSFB_GEMM_HV"); return !e || atoi(e) != 0; }();
  // (GELU layers keep the shared-memory vector path: erff dominates their epilogue and the unrolled form would not fit
  // the instruction cache)
  const bool hv = hv_on && P.n_ct == 1 && P.splits == 1 && (!P.E.bias || host_bias) && (!P.E.norm_w || host_norm_w) &&
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
    FSFB_CUDA(cudaFuncSetAttribute((k_gather_gemm_ts<true, true, false, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget));
    FSFB_CUDA(cudaFuncSetAttribute((k_gather_gemm_ts<true, false, false, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget));
    FSFB_CUDA(cudaFuncSetAttribute((k_gather_gemm_ts<false, false, false, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget));
    FSFB_CUDA(cudaFuncSetAttribute(k_gather_gemm_ts<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget));
    FSFB_CUDA(cudaFuncSetAttribute(k_gather_gemm_ts<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget));
    FSFB_CUDA(cudaFuncSetAttribute(k_gather_gemm_ts<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget));
    attr = true;
  }
  const unsigned grid = (unsigned)std::min<int64_t>(P.n_units, kNumSMs);
  sched_next[dev][P.sched_slot] += (unsigned)P.n_units + grid;
  static const bool timed = [] { const char* e = getenv("FSFB_GEMM_TIMERS"); return e && atoi(e) != 0; }();
  // complete K chunks from 32-byte aligned rows: the 256-bit gather path
  const bool kfull = a_vec && P.cin % kGemmKChunk == 0 && P.cin <= kTsZeroRow && (uintptr_t)P.a % 32 == 0 && P.a_stride % 8 == 0;
  P.timers = nullptr;
  if (timed && kfull) {
    static uint32_t* dev_timers = nullptr;
    if (!dev_timers) {
      FSFB_CUDA(cudaMalloc(&dev_timers, (size_t)kNumSMs * 32 * 4));
      FSFB_CUDA(cudaFuncSetAttribute((k_gather_gemm_ts<true, true, true, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget));
      FSFB_CUDA(cudaFuncSetAttribute((k_gather_gemm_ts<true, true, true, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget));
    }
    FSFB_CUDA(cudaMemsetAsync(dev_timers, 0, (size_t)kNumSMs * 32 * 4, st));
    P.timers = dev_timers;
    g_ts_timers = dev_timers;
    if (hv) {
      FSFB_LAUNCH((k_gather_gemm_ts<true, true, true, true>), grid, kTsThreads, smem, st, P);
    } else {
      FSFB_LAUNCH((k_gather_gemm_ts<true, true, true, false>), grid, kTsThreads, smem, st, P);
    }
  } else if (kfull && gemm_f16_enabled()) {  // experimental fp16-split operands; P.w_packed carries the extra blocks (same switch)
    static bool attr16 = false;
    if (!attr16) {
      FSFB_CUDA(cudaFuncSetAttribute((k_gather_gemm_ts<true, true, false, true, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget));
      FSFB_CUDA(cudaFuncSetAttribute((k_gather_gemm_ts<true, true, false, false, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget));
      attr16 = true;
    }
    if (hv) {
      FSFB_LAUNCH((k_gather_gemm_ts<true, true, false, true, true>), grid, kTsThreads, smem, st, P);
    } else {
      FSFB_LAUNCH((k_gather_gemm_ts<true, true, false, false, true>), grid, kTsThreads, smem, st, P);
    }
  } else if (kfull) {
    if (hv) {
      FSFB_LAUNCH((k_gather_gemm_ts<true, true, false, true>), grid, kTsThreads, smem, st, P);
    } else {
      FSFB_LAUNCH((k_gather_gemm_ts<true, true>), grid, kTsThreads, smem, st, P);
    }
  } else if (a_vec) {
    if (hv) {
      FSFB_LAUNCH((k_gather_gemm_ts<true, false, false, true>), grid, kTsThreads, smem, st, P);
    } else {
      FSFB_LAUNCH((k_gather_gemm_ts<true, false>), grid, kTsThreads, smem, st, P);
    }
  } else {
    if (hv) {
      FSFB_LAUNCH((k_gather_gemm_ts<false, false, false, true>), grid, kTsThreads, smem, st, P);
    } else {
      FSFB_LAUNCH((k_gather_gemm_ts<false, false>), grid, kTsThreads, smem, st, P);
    }
  }
  if (P.splits > 1) return launch_splitk_epilogue(P, st);
  return FSFB_OK;
}

}  // namespace fsfb

// Diagnostics (tools/gemm_timers.py; not part of the product path): copies the per-CTA role counters of the last launch made
// under FSFB_GEMM_TIMERS=1 to host memory: out[148][32] u32.  Layout per CTA: [4g..4g+3] producer group g {wait slot empty,
// convert + tcgen05.st, advance + issue gathers, stages}; [16..19] MMA {open unit, wait accumulators free, wait slot full,
// issue}; [20..24] epilogue {prefetch next unit, wait accumulators full, drain TMEM, finish rows, units}; [31] producer total.
extern "C" int fsfb_debug_gemm_timers(unsigned int* out) {
  using namespace fsfb;
  FSFB_CHECK_ARG(out != nullptr, "debug_gemm_timers: null pointer");
  if (!g_ts_timers) {
    set_error("debug_gemm_timers: no timed launch yet (set FSFB_GEMM_TIMERS=1)");
    return FSFB_ERR_BADARG;
  }
  FSFB_CUDA(cudaDeviceSynchronize());
  FSFB_CUDA(cudaMemcpy(out, g_ts_timers, (size_t)kNumSMs * 32 * 4, cudaMemcpyDeviceToHost));
  return FSFB_OK;
}
