// gemm_lin.cu — dense Linear layers (no neighbour table) over many rows: out = act(norm(a @ w.T + bias) + residual).
//
// Reference call sites: every nn.Linear + LayerNorm/BatchNorm + activation stack the FSF forward builds with build_mlp
// (projects/mmdet3d_plugin/ops/sst_ops.py:808-833) — the pre-voxel encoders, the SIR layers' point MLPs
// (models/backbones/sir.py:41-62), the segmentation / vote heads (models/segmentors/vote_segmentor.py) and the image-feature
// MLPs of FSF.py:138-160.  Same arithmetic as gemm_ss.cu: fp16-split operands (hi + lo / 2048), three kind::f16 tcgen05 MMAs
// per product, main and correction accumulators in tensor memory.
//
// Why a second kernel: the persistent gather kernel (gemm_ss.cu) keeps ONE CTA per SM whose four epilogue warps finish a
// 128 x 128 tile in ~13 k clk while sixteen producer warps share their issue slots; a Linear layer has four K chunks of tensor
// work per tile, so the layer ran at the epilogue's pace (profiles/r2_ss_role_timers.txt: 160 k rows x 128 -> 128 in 107 us,
// 4x its HBM time).  Here a CTA is one 128-row tile and 256 threads that do everything in turn — load + split the A chunk
// (coalesced 128-bit loads, the next three chunks' loads in flight across the barriers), one elected thread issues the MMAs, then all
// eight warps run the epilogue — and two CTAs share an SM (98 KB shared memory, 256 of 512 TMEM columns each), so one tile's
// epilogue overlaps the other tile's loads and MMAs with no hand-written warp specialisation.
//
// Epilogue in two phases over a staging tile that re-uses the operand stages: (1) thread = (row = TMEM lane, column half):
// accumulators + bias -> staging row, LayerNorm statistics of the half (block-wise two-pass, Chan merge) parked in shared memory; (2) every warp walks 16 rows
// with lanes across the columns: per-channel vectors sit in registers, residual reads and output stores are coalesced 128-bit
// accesses.
#include <cuda_fp16.h>

#include <cstdlib>

#include "gemm_persist.cuh"

namespace fsfb {

constexpr int kLinThreads = 256;
constexpr int kLinTile = 128;                       // columns per CTA
constexpr int kLinAStages = 2;
constexpr int kLinWStages = 4;
constexpr uint32_t kLinASlot = kTcRows * 128;       // 16 KB: 128 rows x (32 hi | 32 lo halves)
constexpr uint32_t kLinWSlot = kLinTile * 128;      // 16 KB
constexpr int kLinStageStride = kLinTile + 4;       // floats per staging row (16-byte aligned, conflict-free row walks)
constexpr uint32_t kLinOperandBytes = kLinAStages * kLinASlot + kLinWStages * kLinWSlot;   // 96 KB; the staging tile aliases it
static_assert(kLinOperandBytes >= (uint32_t)kTcRows * kLinStageStride * 4, "staging tile must fit the operand stages");

struct LinShared {
  unsigned long long w_full[kLinWStages];
  unsigned long long w_empty[kLinWStages];
  unsigned long long a_empty[kLinAStages];
  unsigned long long acc_full;
  uint32_t tmem_base;
  alignas(16) float bias[kLinTile];
  float part[2][kTcRows], part2[2][kTcRows];   // per (column half, row): mean and sum of squared deviations of the half's columns
  float mean[kTcRows], rstd[kTcRows];          // merged row statistics (written and read by the row's phase-2 warp)
};

// VEC: rows of `a` are 16-byte aligned with a stride that is a multiple of 4 floats (the stride then covers round_up(cin, 4))
template <bool VEC>
__global__ void __launch_bounds__(kLinThreads, 2) k_linear_ss(const __grid_constant__ TcParams P) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t base = smem_u32(smem_raw);
  const uint32_t s_a = base, s_w = base + kLinAStages * kLinASlot;
  LinShared* sh = reinterpret_cast<LinShared*>(smem_raw + kLinOperandBytes);
  // K split (P.splits > 1, few row tiles x deep K: the 1024-wide refinement heads): CTA z handles the K chunks [kc_lo, kc_hi) and
  // stores raw partial sums to P.partial[z][row][cpad]; k_splitk_epilogue sums the slabs and applies the epilogue
  const bool raw = P.splits > 1;
  Epilogue E = P.E;
  if (raw) {
    E.bias = nullptr;
    E.norm = FSFB_NORM_NONE;
    E.act = FSFB_ACT_NONE;
    E.residual = nullptr;
  }
  const int64_t row0 = (int64_t)blockIdx.x * kTcRows;
  float* const out_base = raw ? P.partial + (int64_t)blockIdx.z * P.rows * P.cpad : P.out;
  const int64_t out_ld = raw ? (int64_t)P.cpad : P.out_stride;
  const int c0 = blockIdx.y * kLinTile;
  const int n_sub = min(kLinTile, P.S.n_pad() - c0);       // MMA N (multiple of 16)
  const int c_n = raw ? kLinTile : min(kLinTile, P.S.cout - c0);   // real channels of this column tile (raw slabs: the padded tile)
  const int kc_all = P.S.kc();
  const int kc_lo = raw ? (int)((int64_t)kc_all * blockIdx.z / P.splits) : 0;
  const int kc_n = (raw ? (int)((int64_t)kc_all * (blockIdx.z + 1) / P.splits) : kc_all) - kc_lo;   // chunks of this CTA: kc_lo + [0, kc_n)
  const uint32_t acc_cols = n_sub <= 64 ? 64u : 128u;

  // ---- operand loads: thread = (16-byte piece of the 128-byte K chunk, rows row_a + 32 i); three chunks in flight ----
  const int chunk = tid & 7, row_a = tid >> 3;
  const bool odd = (chunk & 1) != 0;
  const uint32_t piece = (uint32_t)(odd ? 4 + (chunk >> 1) : (chunk >> 1));   // even lane stores both hi halves, odd both lo
  const uint32_t dst0 = (uint32_t)row_a * 128u + ((piece ^ (uint32_t)(row_a & 7)) << 4);   // + 4096 i: same row & 7
  auto load_chunk = [&](int kc, float4(&v)[4]) {
    const int col = (kc_lo + kc) * kGemmKChunk + chunk * 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int64_t r = row0 + row_a + 32 * i;
      const bool ok = r < P.rows;
      const float* g = P.a + (ok ? r : 0) * P.a_stride + col;
      if (VEC) {
        v[i] = ldg_pred_f4_na(g, ok && col < P.cin);
      } else {
        v[i].x = ldg_pred_f1(g, ok && col < P.cin);
        v[i].y = ldg_pred_f1(g + 1, ok && col + 1 < P.cin);
        v[i].z = ldg_pred_f1(g + 2, ok && col + 2 < P.cin);
        v[i].w = ldg_pred_f1(g + 3, ok && col + 3 < P.cin);
      }
    }
  };
  // the first two chunks are requested before anything else: their latency covers the barrier / tensor-memory set-up below
  float4 q0[4], q1[4], q2[4];   // static register sets: chunk kc lives in q[kc % 3]
  load_chunk(0, q0);
  if (kc_n > 1) load_chunk(1, q1);
  if (kc_n > 2) load_chunk(2, q2);

  // FSFB_GEMM_TIMERS=1 (tools/gemm_role_timers.py): thread 0 of the CTAs 1000..1147 (steady state of a large grid) records
  // the clocks of its phases: [16] prologue, [17] main loop, [18] wait for the last MMA, [19] phase 1, [20] statistics,
  // [21] phase 2, [22] whole CTA
  const bool timed = P.timers != nullptr && tid == 0 && blockIdx.y == 0 && blockIdx.x >= 1000 && blockIdx.x < 1000 + kNumSMs;
  uint32_t* tm = timed ? P.timers + (size_t)(blockIdx.x - 1000) * 32 + 16 : nullptr;
  uint32_t t_last = timed ? (uint32_t)clock() : 0u;
  const uint32_t t_begin = t_last;
#define LIN_T(i) do { if (timed) { const uint32_t t1_ = (uint32_t)clock(); tm[i] = t1_ - t_last; t_last = t1_; } } while (0)
  if (tid == 0) {
    if (base & 1023u) __trap();
    for (int s = 0; s < kLinWStages; ++s) {
      mbar_init(smem_u32(&sh->w_full[s]), 1);
      mbar_init(smem_u32(&sh->w_empty[s]), 1);
    }
    for (int s = 0; s < kLinAStages; ++s) mbar_init(smem_u32(&sh->a_empty[s]), 1);
    mbar_init(smem_u32(&sh->acc_full), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sh->tmem_base)), "r"(2u * acc_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  const float bias_r = (E.bias && tid < c_n) ? __ldg(E.bias + c0 + tid) : 0.f;   // parked in shared memory after the set-up barrier
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = sh->tmem_base;

  // ---- weights: pre-packed (hi | lo) blocks, one bulk copy per K chunk into a ring of four slots ----
  const int nt256 = c0 / kGemmNTile;
  const size_t blk_bytes = P.S.f16_block_bytes(nt256);
  const unsigned char* w_unit = P.w_packed + P.S.f16_tile_base(nt256) + (size_t)(c0 % kGemmNTile) * 128u;
  const uint32_t w_bytes = (uint32_t)n_sub * 128u;
  auto issue_w = [&](int j) {   // thread 0
    const int ws = j & (kLinWStages - 1);
    const uint32_t use = (uint32_t)j / kLinWStages;
    if (use > 0) mbar_wait(smem_u32(&sh->w_empty[ws]), (use - 1u) & 1u);   // the MMAs of chunk j - 4 have read the slot
    mbar_expect_tx(smem_u32(&sh->w_full[ws]), w_bytes);
    bulk_g2s(s_w + (uint32_t)ws * kLinWSlot, w_unit + (size_t)(kc_lo + j) * blk_bytes, w_bytes, smem_u32(&sh->w_full[ws]));
    mbar_arrive(smem_u32(&sh->w_full[ws]));
  };
  if (tid == 0)
    for (int j = 0; j < min(kLinWStages, kc_n); ++j) issue_w(j);

  __half2 ovf = __floats2half2_rn(0.f, 0.f);
  auto store_chunk = [&](int kc, uint32_t slot, const float4(&v)[4]) {
    const int nv = P.cin - ((kc_lo + kc) * kGemmKChunk + chunk * 4);   // real columns in this thread's piece
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float4 x = v[i];
      if (VEC && nv < 4) {   // the 128-bit load ran past cin inside the padded stride: zero what is not input
        if (nv < 2) x.y = 0.f;
        if (nv < 3) x.z = 0.f;
        x.w = 0.f;
        if (nv < 1) x.x = 0.f;
      }
      const __half2 h01 = __floats2half2_rn(x.x, x.y), h23 = __floats2half2_rn(x.z, x.w);
      ovf = __hmax2(ovf, __hmax2(__habs2(h01), __habs2(h23)));
      const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
      const __half2 l01 = __floats2half2_rn((x.x - f01.x) * kF16LoScale, (x.y - f01.y) * kF16LoScale);
      const __half2 l23 = __floats2half2_rn((x.z - f23.x) * kF16LoScale, (x.w - f23.y) * kF16LoScale);
      const uint32_t hi0 = *reinterpret_cast<const uint32_t*>(&h01), hi1 = *reinterpret_cast<const uint32_t*>(&h23);
      const uint32_t lo0 = *reinterpret_cast<const uint32_t*>(&l01), lo1 = *reinterpret_cast<const uint32_t*>(&l23);
      const uint32_t r0 = __shfl_xor_sync(0xffffffffu, odd ? hi0 : lo0, 1);
      const uint32_t r1 = __shfl_xor_sync(0xffffffffu, odd ? hi1 : lo1, 1);
      if (odd) sts_u4(slot + dst0 + 4096u * i, r0, r1, lo0, lo1);
      else sts_u4(slot + dst0 + 4096u * i, hi0, hi1, r0, r1);
    }
  };

  if (tid < kLinTile) sh->bias[tid] = bias_r;   // read in phase 1, behind the main loop's barriers
  LIN_T(0);
  // ---- main loop over the K chunks ----
  const uint32_t idesc = make_idesc_f16(n_sub);
  constexpr uint64_t kDescHi = (uint64_t)0x40004040u << 32;   // SBO = 1024, version 1, SWIZZLE_128B
  auto step = [&](int kc, float4(&v)[4]) {   // v holds chunk kc on entry and chunk kc + 3 on exit
    const int s = kc & 1;
    const uint32_t use = (uint32_t)kc >> 1;   // earlier uses of this A stage
    if (use > 0) mbar_wait(smem_u32(&sh->a_empty[s]), (use - 1u) & 1u);   // the MMAs of chunk kc - 2 have read the stage
    store_chunk(kc, s_a + (uint32_t)s * kLinASlot, v);
    fence_proxy_async();   // generic-proxy stores before the tensor core's async-proxy reads
    if (kc + 3 < kc_n) load_chunk(kc + 3, v);   // in flight across three barriers
    if (tid == 0 && kc + 2 >= kLinWStages && kc + 2 < kc_n) issue_w(kc + 2);
    __syncthreads();
    if (warp == 0) {
      const int ws = kc & (kLinWStages - 1);
      if (lane == 0) mbar_wait(smem_u32(&sh->w_full[ws]), ((uint32_t)kc / kLinWStages) & 1u);
      __syncwarp();
      tc_fence_after();
      uint32_t elected;
      asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(elected));
      if (elected) {
        const uint32_t a_d = (((s_a + (uint32_t)s * kLinASlot) & 0x3FFFFu) >> 4) | 0x10000u;
        const uint32_t w_d = (((s_w + (uint32_t)ws * kLinWSlot) & 0x3FFFFu) >> 4) | 0x10000u;
        const int ksteps = (min(kGemmKChunk, P.cin - (kc_lo + kc) * kGemmKChunk) + 15) >> 4;
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
          if (kk < ksteps) {   // 16-byte units inside the 128-byte row: K step +2, lo half +4
            const uint32_t acc = (kc > 0 || kk > 0) ? 1u : 0u;
            tc_mma_f16_ss(tmem_d, kDescHi | (a_d + 2u * kk), kDescHi | (w_d + 2u * kk), idesc, acc);
            tc_mma_f16_ss(tmem_d + acc_cols, kDescHi | (a_d + 4u + 2u * kk), kDescHi | (w_d + 2u * kk), idesc, acc);
            tc_mma_f16_ss(tmem_d + acc_cols, kDescHi | (a_d + 2u * kk), kDescHi | (w_d + 4u + 2u * kk), idesc, 1u);
          }
        }
        tc_commit(smem_u32(&sh->a_empty[s]));
        tc_commit(smem_u32(&sh->w_empty[ws]));
        if (kc == kc_n - 1) tc_commit(smem_u32(&sh->acc_full));
      }
      __syncwarp();
    }
  };
  for (int kc = 0; kc < kc_n; kc += 3) {
    step(kc, q0);
    if (kc + 1 < kc_n) step(kc + 1, q1);
    if (kc + 2 < kc_n) step(kc + 2, q2);
  }
  {  // overflow report: any converted magnitude that became +Inf
    const __half2 gt = __hgt2(ovf, __floats2half2_rn(65504.f, 65504.f));
    const bool bad = __low2float(gt) != 0.f || __high2float(gt) != 0.f;
    if (__any_sync(0xffffffffu, bad) && lane == 0 && P.ss_overflow) atomicAdd(P.ss_overflow, 1u);
  }

  LIN_T(1);
  // ---- epilogue phase 1: thread = (row = TMEM lane, column half) ----
  mbar_wait(smem_u32(&sh->acc_full), 0u);   // every MMA has completed: the operand stages are free for the staging tile
  tc_fence_after();
  LIN_T(2);
  const int act = E.act & 0xff;
  const bool post = (E.act & FSFB_RESIDUAL_POST) != 0;
  const int quad = warp & 3, half = warp >> 2;
  const int row_l = quad * 32 + lane;
  const int split = n_sub > 64 ? 64 : 32;
  const int cb_lo = half ? split : 0, cb_hi = half ? n_sub : min(split, n_sub);
  const uint32_t t_row = tmem_d + ((uint32_t)(quad * 32) << 16);
  const uint32_t my_row = base + (uint32_t)row_l * (uint32_t)kLinStageStride * 4u;
  const uint32_t s_bias = smem_u32(sh->bias);
  const bool ln = E.norm == FSFB_NORM_LAYERNORM;
  // LayerNorm statistics of this thread's columns: per 32-column block an exact two-pass (mean, sum of squared deviations)
  // on the registers, blocks and halves merged with Chan's update — no second pass over shared memory, no extra barrier
  float cnt = 0.f, mean_h = 0.f, m2_h = 0.f;
#pragma unroll 1
  for (int cb = cb_lo; cb < cb_hi; cb += 32) {   // warp-uniform
    float x[32], c2[32];
    tc_ld32(t_row + cb, x);
    tc_ld32(t_row + acc_cols + cb, c2);
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 bb = lds_f4(s_bias + (uint32_t)(cb + j) * 4u);
      x[j] = fmaf(c2[j], 1.f / kF16LoScale, x[j]) + bb.x;
      x[j + 1] = fmaf(c2[j + 1], 1.f / kF16LoScale, x[j + 1]) + bb.y;
      x[j + 2] = fmaf(c2[j + 2], 1.f / kF16LoScale, x[j + 2]) + bb.z;
      x[j + 3] = fmaf(c2[j + 3], 1.f / kF16LoScale, x[j + 3]) + bb.w;
      // columns in [c_n, n_sub) are exact zeros (zero weight rows, bias 0); past the half's end the 32-column load holds
      // the other half's (or stale) columns
      if (cb + j < cb_hi) sts_f4(my_row + (uint32_t)(cb + j) * 4u, make_float4(x[j], x[j + 1], x[j + 2], x[j + 3]));
    }
    if (ln) {
      const int nb = max(0, min(32, min(c_n, cb_hi) - cb));   // real columns of this block
      if (nb > 0) {
        float sb = 0.f, qb = 0.f, mb;
        if (nb == 32) {   // whole block (every block of a 128-wide layer): no per-column predicates
          float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int j = 0; j < 32; ++j) s4[j & 3] += x[j];
          mb = ((s4[0] + s4[1]) + (s4[2] + s4[3])) * (1.f / 32.f);
          float q4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float d = x[j] - mb;
            q4[j & 3] = fmaf(d, d, q4[j & 3]);
          }
          qb = (q4[0] + q4[1]) + (q4[2] + q4[3]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) sb += j < nb ? x[j] : 0.f;
          mb = sb / (float)nb;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float d = x[j] - mb;
            qb += j < nb ? d * d : 0.f;
          }
        }
        const float tot = cnt + (float)nb, delta = mb - mean_h;
        mean_h += delta * ((float)nb / tot);
        m2_h += qb + delta * delta * (cnt * (float)nb / tot);
        cnt = tot;
      }
    }
  }
  if (raw) {   // the slab is cpad wide: columns past n_sub of this tile are zeros
    const int z_lo = max(half ? split : 0, n_sub), z_hi = half ? kLinTile : split;
    for (int c = z_lo; c < z_hi; c += 4) sts_f4(my_row + (uint32_t)c * 4u, make_float4(0.f, 0.f, 0.f, 0.f));
  }
  tc_fence_before();
  LIN_T(3);
  if (ln) {   // kernel-uniform
    sh->part[half][row_l] = mean_h;
    sh->part2[half][row_l] = m2_h;
  }
  __syncthreads();   // the staging tile (and the row statistics) are complete
  LIN_T(4);

  // ---- epilogue phase 2: warp = 16 rows, lanes across the columns ----
  const bool res_vec = !E.residual || (((uintptr_t)E.residual % 16 == 0) && (E.residual_stride % 4 == 0));
  const bool vec = (c_n & 3) == 0 && P.out_vec && res_vec;
  constexpr int kRowsPerWarp = kTcRows / (kLinThreads / 32);
  const int rl0 = warp * kRowsPerWarp;
  const float n0 = (float)min(c_n, split), n1 = (float)(c_n - min(c_n, split));
  // lane i < 16 merges the two halves' (mean, M2) of row rl0 + i once; the rows are this warp's own, so a warp barrier orders
  // the broadcast reads of the row walk below
  if (ln) {
    if (lane < kRowsPerWarp) {
      const int rl = rl0 + lane;
      const float m0 = sh->part[0][rl], m1 = sh->part[1][rl];
      const float delta = m1 - m0;
      sh->mean[rl] = m0 + delta * (n1 / (float)c_n);
      const float m2 = sh->part2[0][rl] + sh->part2[1][rl] + delta * delta * (n0 * n1 / (float)c_n);
      sh->rstd[rl] = 1.f / sqrtf(m2 / (float)c_n + E.eps);
    }
    __syncwarp();
  }
  if (vec) {
    const int c = 4 * lane;
    const bool on = c < c_n;
    float4 w4 = make_float4(1.f, 1.f, 1.f, 1.f), h4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (on && E.norm != FSFB_NORM_NONE) {
      w4 = __ldg(reinterpret_cast<const float4*>(E.norm_w + c0 + c));
      h4 = __ldg(reinterpret_cast<const float4*>(E.norm_b + c0 + c));
    }
    if (on) {
#pragma unroll 4
      for (int i = 0; i < kRowsPerWarp; ++i) {
        const int64_t r = row0 + rl0 + i;
        if (r >= P.rows) break;
        float4 y = lds_f4(base + (uint32_t)((rl0 + i) * kLinStageStride + c) * 4u);
        if (ln) {
          const float m = sh->mean[rl0 + i], rs = sh->rstd[rl0 + i];
          y.x = (y.x - m) * rs * w4.x + h4.x; y.y = (y.y - m) * rs * w4.y + h4.y;
          y.z = (y.z - m) * rs * w4.z + h4.z; y.w = (y.w - m) * rs * w4.w + h4.w;
        } else if (E.norm == FSFB_NORM_AFFINE) {
          y.x = fmaf(y.x, w4.x, h4.x); y.y = fmaf(y.y, w4.y, h4.y); y.z = fmaf(y.z, w4.z, h4.z); y.w = fmaf(y.w, w4.w, h4.w);
        }
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        if (E.residual) g = __ldg(reinterpret_cast<const float4*>(E.residual + r * E.residual_stride + c0 + c));
        if (post) {
          y.x = apply_act(y.x, act) + g.x; y.y = apply_act(y.y, act) + g.y; y.z = apply_act(y.z, act) + g.z; y.w = apply_act(y.w, act) + g.w;
        } else {
          y.x = apply_act(y.x + g.x, act); y.y = apply_act(y.y + g.y, act); y.z = apply_act(y.z + g.z, act); y.w = apply_act(y.w + g.w, act);
        }
        *reinterpret_cast<float4*>(out_base + r * out_ld + c0 + c) = y;
      }
    }
  } else {   // odd widths / unaligned rows: the same walk, one column per lane and trip
    float w1[kLinTile / 32], h1[kLinTile / 32];
#pragma unroll
    for (int j = 0; j < kLinTile / 32; ++j) {
      const int c = lane + 32 * j;
      const bool on = c < c_n && E.norm != FSFB_NORM_NONE;
      w1[j] = on ? __ldg(E.norm_w + c0 + c) : 1.f;
      h1[j] = on ? __ldg(E.norm_b + c0 + c) : 0.f;
    }
    for (int i = 0; i < kRowsPerWarp; ++i) {
      const int64_t r = row0 + rl0 + i;
      if (r >= P.rows) break;   // warp-uniform
      const float m = ln ? sh->mean[rl0 + i] : 0.f, rs = ln ? sh->rstd[rl0 + i] : 1.f;
#pragma unroll
      for (int j = 0; j < kLinTile / 32; ++j) {
        const int c = lane + 32 * j;
        if (c < c_n) {
          float y;
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(y) : "r"(base + (uint32_t)((rl0 + i) * kLinStageStride + c) * 4u));
          if (ln) y = (y - m) * rs * w1[j] + h1[j];
          else if (E.norm == FSFB_NORM_AFFINE) y = fmaf(y, w1[j], h1[j]);
          const float g = E.residual ? __ldg(E.residual + r * E.residual_stride + c0 + c) : 0.f;
          y = post ? apply_act(y, act) + g : apply_act(y + g, act);
          out_base[r * out_ld + c0 + c] = y;
        }
      }
    }
  }
  __syncthreads();   // every warp is done with tensor memory
  LIN_T(5);
  if (timed) tm[6] = (uint32_t)clock() - t_begin;
#undef LIN_T
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(2u * acc_cols) : "memory");
  }
}

// 0 = launched, 1 = shape not served here (the persistent kernels take it), otherwise an error code
int launch_linear_ss(TcParams& P, bool a_vec, int ksplits, float* workspace, size_t workspace_bytes, cudaStream_t st) {
  static const int mode = [] { const char* e = getenv("FSFB_GEMM_LIN"); return e ? atoi(e) : 1; }();
  if (!mode || !gemm_f16_enabled()) return 1;
  static const int64_t min_rows = [] { const char* e = getenv("FSFB_GEMM_LIN_MIN_ROWS"); return e ? atoll(e) : 1024ll; }();
  const int n_pad = P.S.n_pad();
  // below min_rows only narrow layers come here (cout <= 256: a handful of CTAs beat the persistent kernel's fixed cost — 43 -> 20 us
  // for 248 x 1024 -> 128 — while 1024-wide outputs over two row tiles do not)
  if (P.nbr || P.row_order || P.koff != 1 || (P.rows < min_rows && n_pad > 256)) return 1;
  if (n_pad > kLinTile && P.E.norm == FSFB_NORM_LAYERNORM && ksplits <= 1) return 1;   // row statistics across column tiles
  if (P.a_rows < P.rows) return 1;
  {
    const int rc = ss_overflow_counter(&P.ss_overflow);
    if (rc != FSFB_OK) return rc;
  }
  {
    const int rc = ss_timers_buffer(&P.timers, st);
    if (rc != FSFB_OK) return rc;
  }
  const size_t smem = (size_t)kLinOperandBytes + sizeof(LinShared);
  static bool attr = false;
  if (!attr) {
    FSFB_CUDA(cudaFuncSetAttribute(k_linear_ss<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    FSFB_CUDA(cudaFuncSetAttribute(k_linear_ss<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  // K split: raw slabs [ksplits][rows][cpad] + the split epilogue kernel (bias / norm / activation / residual)
  P.splits = 1;
  P.partial = nullptr;
  P.cpad = (n_pad + 127) & ~127;
  if (ksplits > 1) {
    const size_t need = (size_t)ksplits * (size_t)P.rows * P.cpad * sizeof(float);
    if (ksplits > P.S.kc() || P.cpad > 1024 || !workspace || workspace_bytes < need) {
      set_error("gather_gemm: K split %d needs 1 < splits <= %d K chunks, cout <= 1024 and %zu bytes of workspace (%zu given)", ksplits,
                P.S.kc(), need, workspace_bytes);
      return FSFB_ERR_CAPACITY;
    }
    P.splits = ksplits;
    P.partial = workspace;
  }
  const dim3 grid((unsigned)ceil_div(P.rows, kTcRows), (unsigned)ceil_div(n_pad, kLinTile), (unsigned)P.splits);
  if (a_vec) FSFB_LAUNCH(k_linear_ss<true>, grid, kLinThreads, smem, st, P);
  else FSFB_LAUNCH(k_linear_ss<false>, grid, kLinThreads, smem, st, P);
  if (P.splits > 1) return launch_splitk_epilogue(P, st);
  return FSFB_OK;
}

}  // namespace fsfb
