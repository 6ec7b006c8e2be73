// segment_reduce.cu — segmented max(+argmax) / mean / sum over rows grouped by a CSR (a2, a12, a13)
// and the row gather that undoes it (a6).
//
// Reference semantics: torch_scatter.scatter_max(feat, inv, dim=0) and
// torch_scatter.scatter(feat, inv, dim=0, reduce='mean'|'sum') as called from scatter_v2,
// projects/mmdet3d_plugin/ops/sst_ops.py:168,170.  argmax = lowest source row attaining the max
// (torch_scatter's sequential CPU rule); empty segment → value 0, argmax = n.
//
// B200 design (HBM-bound; algorithmic bytes 4NC + 8N + 4MC): no atomics on features.  Rows are
// visited in CSR order; a warp owns 32 consecutive sorted positions, lanes own channels
// (128-bit loads when C % 4 == 0), rows stream through registers 4 at a time, and a segment
// that begins and ends inside the warp's chunk is written once with plain stores.  Segments
// that cross a chunk boundary leave at most two partial rows per chunk in a scratch buffer;
// a second small kernel folds those (and zero-fills empty segments).
#include <cfloat>

#include "common.cuh"

namespace fsfb {

constexpr int kSrChunk = 32;    // sorted positions per warp
constexpr int kSrWarps = 8;     // warps per CTA
constexpr int kSrUnroll = 4;    // rows in flight per warp

template <int VEC>
__device__ __forceinline__ void load_row(const float* __restrict__ p, bool ok, float (&v)[VEC]) {
  if (VEC == 4) {
    float4 t = ok ? ldg_stream_f4(reinterpret_cast<const float4*>(p)) : make_float4(0, 0, 0, 0);
    v[0] = t.x;
    v[VEC > 1 ? 1 : 0] = t.y;
    v[VEC > 2 ? 2 : 0] = t.z;
    v[VEC > 3 ? 3 : 0] = t.w;
  } else {
    v[0] = ok ? ldg_stream_f1(p) : 0.f;
  }
}

template <int VEC, int K, bool IS_MAX, bool HAS_ARG>
struct Acc {
  float val[K][VEC];
  int arg[HAS_ARG ? K : 1][HAS_ARG ? VEC : 1];
  __device__ __forceinline__ void reset() {
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        val[k][e] = IS_MAX ? -INFINITY : 0.f;
        if (HAS_ARG) arg[k][e] = INT_MAX;
      }
  }
  __device__ __forceinline__ void add(int k, const float (&v)[VEC], int row) {
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      if (IS_MAX) {
        if (HAS_ARG) {
          bool take = (v[e] > val[k][e]) | ((v[e] == val[k][e]) & (row < arg[k][e]));
          arg[k][e] = take ? row : arg[k][e];
          val[k][e] = take ? v[e] : val[k][e];
        } else {
          val[k][e] = fmaxf(val[k][e], v[e]);
        }
      } else {
        val[k][e] += v[e];
      }
    }
  }
};

// Channel owned by (lane, k, e) inside the channel block starting at c0.
template <int VEC>
__device__ __forceinline__ int chan(int c0, int lane, int k, int e) {
  return c0 + (k * 32 + lane) * VEC + e;
}

template <int VEC, int K, bool IS_MAX, bool HAS_ARG>
__global__ void __launch_bounds__(kSrWarps * 32)
    k_segreduce(const float* __restrict__ feat, int64_t stride, int C,
                const int32_t* __restrict__ perm, const int32_t* __restrict__ seg,
                const int32_t* __restrict__ offsets, int m, int mean, float* __restrict__ out,
                long long* __restrict__ argout, int32_t* __restrict__ part_seg,
                float* __restrict__ part_val, int32_t* __restrict__ part_arg, int64_t n_chunks) {
  const int lane = lane_id();
  const int64_t gw = (int64_t)blockIdx.x * kSrWarps + (threadIdx.x >> 5);
  if (gw >= n_chunks) return;  // warp-uniform
  const int c0 = blockIdx.y * (32 * VEC * K);
  const int64_t n_valid = offsets[m];
  const int64_t p0 = gw * kSrChunk;
  const int cnt = (int)max((int64_t)0, min((int64_t)kSrChunk, n_valid - p0));
  int head_seg = -1, tail_seg = -1;  // what lands in the two partial slots
  if (cnt > 0) {
    const int64_t p = p0 + lane;
    const int my_seg = lane < cnt ? seg[p] : -1;
    const int my_row = lane < cnt ? (perm ? perm[p] : (int)p) : 0;
    const int seg_prev = p0 > 0 ? seg[p0 - 1] : -1;
    const int seg_next = (p0 + kSrChunk < n_valid) ? seg[p0 + kSrChunk] : -2;

    Acc<VEC, K, IS_MAX, HAS_ARG> acc;
    acc.reset();
    int cur = __shfl_sync(0xffffffffu, my_seg, 0);
    int run = 0;  // rows accumulated into `cur` inside this chunk (== segment size when it is not partial)

    auto flush = [&](int s) {
      const bool partial = (s == seg_prev) | (s == seg_next);
      if (!partial) {
        const int cnt_r = max(1, run);
        const bool pow2 = (cnt_r & (cnt_r - 1)) == 0;  // exact multiply instead of the division sequence (same bits)
        const float denom = mean ? (float)cnt_r : 1.f;
        const float recip = 1.f / denom;
#pragma unroll
        for (int k = 0; k < K; ++k) {
          const int c = chan<VEC>(c0, lane, k, 0);
          if (c < C) {
            float r[VEC];
#pragma unroll
            for (int e = 0; e < VEC; ++e) r[e] = mean ? (pow2 ? acc.val[k][e] * recip : acc.val[k][e] / denom) : acc.val[k][e];
            float* o = out + (int64_t)s * C + c;
            if (VEC == 4) {
              stg_stream_f4(reinterpret_cast<float4*>(o),
                            make_float4(r[0], r[VEC > 1 ? 1 : 0], r[VEC > 2 ? 2 : 0], r[VEC > 3 ? 3 : 0]));
            } else {
              o[0] = r[0];
            }
            if (HAS_ARG) {
#pragma unroll
              for (int e = 0; e < VEC; ++e) argout[(int64_t)s * C + c + e] = (long long)acc.arg[k][e];
            }
          }
        }
      } else {
        const int64_t slot = 2 * gw + ((s == seg_prev) ? 0 : 1);
        if (s == seg_prev) head_seg = s; else tail_seg = s;
#pragma unroll
        for (int k = 0; k < K; ++k) {
          const int c = chan<VEC>(c0, lane, k, 0);
          if (c < C) {
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
              part_val[slot * C + c + e] = acc.val[k][e];
              if (HAS_ARG) part_arg[slot * C + c + e] = acc.arg[k][e];
            }
          }
        }
      }
      acc.reset();
      run = 0;
    };

    for (int r0 = 0; r0 < cnt; r0 += kSrUnroll) {
      float v[kSrUnroll][K][VEC];
      int rows[kSrUnroll], segs[kSrUnroll];
#pragma unroll
      for (int u = 0; u < kSrUnroll; ++u) {
        const int r = min(r0 + u, kSrChunk - 1);
        rows[u] = __shfl_sync(0xffffffffu, my_row, r);
        segs[u] = __shfl_sync(0xffffffffu, my_seg, r);
        const bool live = (r0 + u) < cnt;
        const float* base = feat + (int64_t)rows[u] * stride;
#pragma unroll
        for (int k = 0; k < K; ++k) {
          const int c = chan<VEC>(c0, lane, k, 0);
          load_row<VEC>(base + c, live && c < C, v[u][k]);
        }
      }
#pragma unroll
      for (int u = 0; u < kSrUnroll; ++u) {
        if (r0 + u < cnt) {
          if (segs[u] != cur) {
            flush(cur);
            cur = segs[u];
          }
#pragma unroll
          for (int k = 0; k < K; ++k) acc.add(k, v[u][k], rows[u]);
          ++run;
        }
      }
    }
    flush(cur);
  }
  if (lane == 0 && blockIdx.y == 0) {
    part_seg[2 * gw] = head_seg;
    part_seg[2 * gw + 1] = tail_seg;
  }
}

// ---- short segments: one warp per OUTPUT segment ---------------------------------------------------
// Voxel-level reductions (0.1-0.2 m voxels hold 1-2 points on average) are dominated by the chunk kernel's
// boundary bookkeeping; here a warp owns a segment, walks its rows two at a time and writes the result once —
// no partial slots, no fix-up pass, no workspace.  Used when n / m <= kSrSmallAvg.
constexpr int kSrSmallAvg = 8;

// A warp handles PAIRS of adjacent segments and runs a three-level software pipeline over its pairs, because the walk
// offsets[s] → perm[p] → feat[row] is three dependent global loads: while the rows of pair i are in flight, the row ids
// of pair i+1 and the offsets of pair i+2 are already requested (each is one value per lane).  Measured before: one
// segment at a time left the kernel latency-bound at 42 % of HBM peak on the 300 k x 132 pre-voxel mean.
template <int VEC, int K, bool IS_MAX, bool HAS_ARG>
__global__ void __launch_bounds__(kSrWarps * 32)
    k_segreduce_small(const float* __restrict__ feat, int64_t stride, int C, const int32_t* __restrict__ perm,
                      const int32_t* __restrict__ offsets, int m, int mean, int64_t n, float* __restrict__ out,
                      long long* __restrict__ argout) {
  const int lane = lane_id();
  const int c0 = blockIdx.y * (32 * VEC * K);
  const int64_t n_warps = (int64_t)gridDim.x * kSrWarps;
  const int64_t w = (int64_t)blockIdx.x * kSrWarps + (threadIdx.x >> 5);
  auto pair_base = [&](int64_t it) { return 2 * (w + it * n_warps); };
  // level 1: lanes 0..2 hold offsets[s], offsets[s+1], offsets[s+2] of the pair starting at segment s
  auto load_off = [&](int64_t s) -> int {
    int v = 0;
    if (lane < 3 && s + lane <= m) v = __ldg(offsets + s + lane);
    return v;
  };
  // level 2: lanes 0..3 hold the first two row ids of both segments (-1: the segment has no such row)
  auto load_perm = [&](int off, int64_t s) -> int {
    const int b0 = __shfl_sync(0xffffffffu, off, 0), b1 = __shfl_sync(0xffffffffu, off, 1);
    const int e1 = s + 1 < m ? __shfl_sync(0xffffffffu, off, 2) : b1;
    int r = -1;
    if (lane < 4 && s < m) {
      const int p = (lane < 2 ? b0 : b1) + (lane & 1);
      if (p < (lane < 2 ? b1 : e1)) r = perm ? __ldg(perm + p) : p;
    }
    return r;
  };
  int off_a = load_off(pair_base(0));
  int off_b = load_off(pair_base(1));
  int prm_a = load_perm(off_a, pair_base(0));
  for (int64_t it = 0;; ++it) {
    const int64_t s = pair_base(it);
    if (s >= m) break;
    const int off_c = load_off(pair_base(it + 2));
    const int prm_b = load_perm(off_b, pair_base(it + 1));
    int beg[2], end[2], row[2][2];
    beg[0] = __shfl_sync(0xffffffffu, off_a, 0);
    end[0] = beg[1] = __shfl_sync(0xffffffffu, off_a, 1);
    end[1] = s + 1 < m ? __shfl_sync(0xffffffffu, off_a, 2) : end[0];
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int i = 0; i < 2; ++i) row[j][i] = __shfl_sync(0xffffffffu, prm_a, 2 * j + i);
    // level 3: up to four rows in flight
    float v[2][2][K][VEC];
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int k = 0; k < K; ++k) {
          const int c = chan<VEC>(c0, lane, k, 0);
          load_row<VEC>(feat + (int64_t)max(row[j][i], 0) * stride + c, row[j][i] >= 0 && c < C, v[j][i][k]);
        }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      if (s + j >= m) break;  // warp-uniform
      Acc<VEC, K, IS_MAX, HAS_ARG> acc;
      acc.reset();
#pragma unroll
      for (int i = 0; i < 2; ++i)
        if (row[j][i] >= 0) {
#pragma unroll
          for (int k = 0; k < K; ++k) acc.add(k, v[j][i][k], row[j][i]);
        }
      for (int p = beg[j] + 2; p < end[j]; p += 2) {  // longer segments: the rest, two rows at a time
        const bool two = p + 1 < end[j];
        const int r0 = perm ? __ldg(perm + p) : p;
        const int r1 = two ? (perm ? __ldg(perm + p + 1) : p + 1) : r0;
        float v0[K][VEC], v1[K][VEC];
#pragma unroll
        for (int k = 0; k < K; ++k) {
          const int c = chan<VEC>(c0, lane, k, 0);
          load_row<VEC>(feat + (int64_t)r0 * stride + c, c < C, v0[k]);
          load_row<VEC>(feat + (int64_t)r1 * stride + c, two && c < C, v1[k]);
        }
#pragma unroll
        for (int k = 0; k < K; ++k) acc.add(k, v0[k], r0);
        if (two) {
#pragma unroll
          for (int k = 0; k < K; ++k) acc.add(k, v1[k], r1);
        }
      }
      const bool empty = end[j] == beg[j];
      // mean = sum / count in IEEE fp32 (what the reference computes); counts that are powers of two (1 and 2 cover most
      // voxels) take an exact multiply instead of the ~8-instruction division sequence: same bits, fewer instructions
      const int cnt_j = max(1, end[j] - beg[j]);
      const bool pow2 = (cnt_j & (cnt_j - 1)) == 0;
      const float denom = mean ? (float)cnt_j : 1.f;
      const float recip = 1.f / denom;  // exact when pow2
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const int c = chan<VEC>(c0, lane, k, 0);
        if (c < C) {
          float r[VEC];
#pragma unroll
          for (int e = 0; e < VEC; ++e)
            r[e] = empty ? 0.f : (mean ? (pow2 ? acc.val[k][e] * recip : acc.val[k][e] / denom) : acc.val[k][e]);
          float* o = out + (s + j) * C + c;
          if (VEC == 4) {
            stg_stream_f4(reinterpret_cast<float4*>(o), make_float4(r[0], r[VEC > 1 ? 1 : 0], r[VEC > 2 ? 2 : 0], r[VEC > 3 ? 3 : 0]));
          } else {
            o[0] = r[0];
          }
          if (HAS_ARG) {
#pragma unroll
            for (int e = 0; e < VEC; ++e) argout[(s + j) * C + c + e] = empty ? (long long)n : (long long)acc.arg[k][e];
          }
        }
      }
    }
    off_a = off_b;
    off_b = off_c;
    prm_a = prm_b;
  }
}

// C == 132 (the 131-channel point features in their 16-byte padded rows: the widest scatter of the frame).  The generic
// kernel would run it as K = 2 with one live lane in its second half: twice the load instructions and 70 registers.
// Here the warp covers float4 0..31 of a row as above and the 33rd float4 of the (up to) four rows in flight is fetched by
// lanes 0..3 in ONE instruction — lane 2j+i holds row i of segment j — and folded with a single shuffle step.
template <bool IS_MAX>
__global__ void __launch_bounds__(kSrWarps * 32)
    k_segreduce_small_132(const float* __restrict__ feat, int64_t stride, const int32_t* __restrict__ perm,
                          const int32_t* __restrict__ offsets, int m, int mean, float* __restrict__ out) {
  constexpr int C = 132;
  const int lane = lane_id();
  const int64_t n_warps = (int64_t)gridDim.x * kSrWarps;
  const int64_t w = (int64_t)blockIdx.x * kSrWarps + (threadIdx.x >> 5);
  auto pair_base = [&](int64_t it) { return 2 * (w + it * n_warps); };
  auto load_off = [&](int64_t s) -> int {
    int v = 0;
    if (lane < 3 && s + lane <= m) v = __ldg(offsets + s + lane);
    return v;
  };
  auto load_perm = [&](int off, int64_t s) -> int {
    const int b0 = __shfl_sync(0xffffffffu, off, 0), b1 = __shfl_sync(0xffffffffu, off, 1);
    const int e1 = s + 1 < m ? __shfl_sync(0xffffffffu, off, 2) : b1;
    int r = -1;
    if (lane < 4 && s < m) {
      const int p = (lane < 2 ? b0 : b1) + (lane & 1);
      if (p < (lane < 2 ? b1 : e1)) r = perm ? __ldg(perm + p) : p;
    }
    return r;
  };
  auto comb = [&](float a, float b) { return IS_MAX ? fmaxf(a, b) : a + b; };
  const float ident = IS_MAX ? -INFINITY : 0.f;
  int off_a = load_off(pair_base(0));
  int off_b = load_off(pair_base(1));
  int prm_a = load_perm(off_a, pair_base(0));
  for (int64_t it = 0;; ++it) {
    const int64_t s = pair_base(it);
    if (s >= m) break;
    const int off_c = load_off(pair_base(it + 2));
    const int prm_b = load_perm(off_b, pair_base(it + 1));
    int beg[2], end[2], row[2][2];
    beg[0] = __shfl_sync(0xffffffffu, off_a, 0);
    end[0] = beg[1] = __shfl_sync(0xffffffffu, off_a, 1);
    end[1] = s + 1 < m ? __shfl_sync(0xffffffffu, off_a, 2) : end[0];
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int i = 0; i < 2; ++i) row[j][i] = __shfl_sync(0xffffffffu, prm_a, 2 * j + i);
    float4 v[2][2];
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        v[j][i] = make_float4(ident, ident, ident, ident);
        if (row[j][i] >= 0) v[j][i] = ldg_stream_f4(reinterpret_cast<const float4*>(feat + (int64_t)row[j][i] * stride) + lane);
      }
    // 33rd float4: lane l < 4 owns row (l >> 1, l & 1); prm_a of that lane is exactly its row id
    float4 tv = make_float4(ident, ident, ident, ident);
    if (lane < 4 && prm_a >= 0) tv = ldg_stream_f4(reinterpret_cast<const float4*>(feat + (int64_t)prm_a * stride) + 32);
    {
      float4 o;
      o.x = __shfl_xor_sync(0xffffffffu, tv.x, 1); o.y = __shfl_xor_sync(0xffffffffu, tv.y, 1);
      o.z = __shfl_xor_sync(0xffffffffu, tv.z, 1); o.w = __shfl_xor_sync(0xffffffffu, tv.w, 1);
      tv = make_float4(comb(tv.x, o.x), comb(tv.y, o.y), comb(tv.z, o.z), comb(tv.w, o.w));  // lanes 0 / 2: segments 0 / 1
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      if (s + j >= m) break;  // warp-uniform
      float4 a = make_float4(comb(v[j][0].x, v[j][1].x), comb(v[j][0].y, v[j][1].y), comb(v[j][0].z, v[j][1].z), comb(v[j][0].w, v[j][1].w));
      float4 t = make_float4(__shfl_sync(0xffffffffu, tv.x, 2 * j), __shfl_sync(0xffffffffu, tv.y, 2 * j),
                             __shfl_sync(0xffffffffu, tv.z, 2 * j), __shfl_sync(0xffffffffu, tv.w, 2 * j));
      for (int p = beg[j] + 2; p < end[j]; ++p) {  // longer segments: the rest, one row at a time
        const int r = perm ? __ldg(perm + p) : p;
        const float4 x = ldg_stream_f4(reinterpret_cast<const float4*>(feat + (int64_t)r * stride) + lane);
        const float4 y = ldg_stream_f4(reinterpret_cast<const float4*>(feat + (int64_t)r * stride) + 32);  // same address in every lane
        a = make_float4(comb(a.x, x.x), comb(a.y, x.y), comb(a.z, x.z), comb(a.w, x.w));
        t = make_float4(comb(t.x, y.x), comb(t.y, y.y), comb(t.z, y.z), comb(t.w, y.w));
      }
      const bool empty = end[j] == beg[j];
      const int cnt_j = max(1, end[j] - beg[j]);
      const bool pow2 = (cnt_j & (cnt_j - 1)) == 0;
      const float denom = mean ? (float)cnt_j : 1.f;
      const float recip = 1.f / denom;
      auto fin = [&](float x) { return empty ? 0.f : (mean ? (pow2 ? x * recip : x / denom) : x); };
      float* o = out + (s + j) * C;
      stg_stream_f4(reinterpret_cast<float4*>(o) + lane, make_float4(fin(a.x), fin(a.y), fin(a.z), fin(a.w)));
      if (lane == 0) stg_stream_f4(reinterpret_cast<float4*>(o) + 32, make_float4(fin(t.x), fin(t.y), fin(t.z), fin(t.w)));
    }
    off_a = off_b;
    off_b = off_c;
    prm_a = prm_b;
  }
}

// Fold partial rows of segments that cross chunk boundaries; zero-fill empty segments.
// One warp per (chunk, block of 32 channels); the walk over the chunks a long segment spans is unrolled
// four-wide so its loads are independent (instance-level segments can span hundreds of chunks).
template <bool IS_MAX, bool HAS_ARG>
__global__ void __launch_bounds__(kSrWarps * 32)
    k_segreduce_fixup(int C, const int32_t* __restrict__ offsets, int m, int mean, int dense, int64_t n,
                      float* __restrict__ out, long long* __restrict__ argout,
                      const int32_t* __restrict__ part_seg, const float* __restrict__ part_val,
                      const int32_t* __restrict__ part_arg, int64_t n_chunks) {
  const int lane = lane_id();
  const int cblocks = (C + 31) / 32;
  const int64_t gwarp = (int64_t)blockIdx.x * kSrWarps + (threadIdx.x >> 5);
  const int64_t gw = gwarp / cblocks;
  const int c = (int)(gwarp - gw * cblocks) * 32 + lane;
  // part 1: owner = chunk whose tail slot holds a segment that started inside it
  if (gw < n_chunks) {
    const int s = part_seg[2 * gw + 1];
    if (s >= 0 && c < C) {
      const int64_t last_chunk = ((int64_t)offsets[s + 1] - 1) / kSrChunk;
      const int count = offsets[s + 1] - offsets[s];
      float best = part_val[(2 * gw + 1) * C + c];
      int arg = HAS_ARG ? part_arg[(2 * gw + 1) * C + c] : 0;
      auto fold = [&](float v, int a) {
        if (IS_MAX) {
          if (HAS_ARG) {
            const bool take = (v > best) | ((v == best) & (a < arg));
            arg = take ? a : arg;
            best = take ? v : best;
          } else {
            best = fmaxf(best, v);
          }
        } else {
          best += v;
        }
      };
      int64_t w = gw + 1;
      for (; w + 3 <= last_chunk; w += 4) {
        float v[4];
        int a[4] = {0, 0, 0, 0};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          v[u] = part_val[(2 * (w + u)) * C + c];
          if (HAS_ARG) a[u] = part_arg[(2 * (w + u)) * C + c];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) fold(v[u], a[u]);  // ascending chunk order: deterministic sums
      }
      for (; w <= last_chunk; ++w) fold(part_val[(2 * w) * C + c], HAS_ARG ? part_arg[(2 * w) * C + c] : 0);
      if (mean) best = best / (float)max(1, count);
      out[(int64_t)s * C + c] = best;
      if (HAS_ARG) argout[(int64_t)s * C + c] = (long long)arg;
    }
  }
  // part 2: empty segments (only possible when the index does not come from a ranking: `dense` skips it)
  if (dense) return;
  const int64_t n_warps = (int64_t)gridDim.x * kSrWarps;
  for (int64_t s = gwarp; s < m; s += n_warps) {
    if (offsets[s + 1] == offsets[s]) {
      for (int cc = lane; cc < C; cc += 32) {
        out[s * C + cc] = 0.f;
        if (HAS_ARG) argout[s * C + cc] = (long long)n;
      }
    }
  }
}

// ---- row gather ------------------------------------------------------------------------
// Flat mapping: one thread per 16-byte (VEC = 4) or 4-byte chunk of the output, so every lane is busy
// whatever the row width (C = 3 ... 768 on this path); 4 chunks in flight per thread.
template <typename IdxT, int VEC>
__global__ void __launch_bounds__(256)
    k_gather_rows(const float* __restrict__ src, int64_t m, int C, int64_t src_stride, const IdxT* __restrict__ idx,
                  int64_t n, float fill, float* __restrict__ out, int64_t out_stride) {
  const int cv = C / VEC;
  const int64_t total = n * cv;
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t0 < total; t0 += 4 * step) {
    float4 v[4];
    int64_t dst[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t t = t0 + u * step;
      dst[u] = -1;
      if (t < total) {
        const int64_t i = t / cv;
        const int j = (int)(t - i * cv) * VEC;
        const long long s = (long long)__ldg(idx + i);
        const bool ok = s >= 0 && s < m;
        dst[u] = i * out_stride + j;
        if (VEC == 4) {
          v[u] = ok ? __ldg(reinterpret_cast<const float4*>(src + s * src_stride + j)) : make_float4(fill, fill, fill, fill);
        } else {
          v[u].x = ok ? __ldg(src + s * src_stride + j) : fill;
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (dst[u] >= 0) {
        if (VEC == 4) stg_stream_f4(reinterpret_cast<float4*>(out + dst[u]), v[u]);
        else out[dst[u]] = v[u].x;
      }
    }
  }
}

template <int VEC, int K>
static int launch_segreduce_small(const float* feat, int64_t stride, int C, const int32_t* perm, const int32_t* offsets,
                                  int m, int mode, int64_t n, float* out, long long* argout, cudaStream_t st) {
  const int mean = mode == FSFB_REDUCE_MAX ? 0 : (mode == FSFB_REDUCE_MEAN);
  auto kern = (mode == FSFB_REDUCE_MAX)
                  ? (argout ? k_segreduce_small<VEC, K, true, true> : k_segreduce_small<VEC, K, true, false>)
                  : k_segreduce_small<VEC, K, false, false>;
  // persistent grid-stride warps: one wave of resident CTAs (the software pipeline wants long runs per warp)
  static int resident_of[3] = {0, 0, 0};  // per instantiation of this template: max / max+arg / sum-mean kernels
  int& resident = resident_of[mode == FSFB_REDUCE_MAX ? (argout ? 1 : 0) : 2];
  if (resident == 0) FSFB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, kern, kSrWarps * 32, 0));
  const int cblocks = (int)ceil_div(C, 32 * VEC * K);
  const int64_t wave = (int64_t)kNumSMs * std::max(1, resident / cblocks);
  dim3 grid((unsigned)std::min<int64_t>(ceil_div(ceil_div(m, 2), kSrWarps), wave), (unsigned)cblocks);
  FSFB_LAUNCH(kern, grid, kSrWarps * 32, 0, st, feat, stride, C, perm, offsets, m, mean, n, out, argout);
  return FSFB_OK;
}

template <int VEC, int K>
static int launch_segreduce(const float* feat, int64_t stride, int C, const int32_t* perm,
                            const int32_t* seg, const int32_t* offsets, int m, int mode, float* out,
                            long long* argout, int32_t* part_seg, float* part_val,
                            int32_t* part_arg, int64_t n_chunks, cudaStream_t st) {
  dim3 grid((unsigned)ceil_div(n_chunks, kSrWarps), (unsigned)ceil_div(C, 32 * VEC * K));
  const int mean = mode == FSFB_REDUCE_MEAN;
  auto kern = (mode == FSFB_REDUCE_MAX)
                  ? (argout ? k_segreduce<VEC, K, true, true> : k_segreduce<VEC, K, true, false>)
                  : k_segreduce<VEC, K, false, false>;
  FSFB_LAUNCH(kern, grid, kSrWarps * 32, 0, st, feat, stride, C, perm, seg, offsets, m, mean, out,
              argout, part_seg, part_val, part_arg, n_chunks);
  return FSFB_OK;
}

}  // namespace fsfb

extern "C" {

int fsfb_segment_reduce_workspace_bytes(int64_t n, int c, int with_arg, size_t* bytes) {
  using namespace fsfb;
  FSFB_CHECK_ARG(bytes && n >= 0 && c >= 1, "segment_reduce_workspace_bytes: bad argument");
  const int64_t n_chunks = std::max<int64_t>(1, ceil_div(n, kSrChunk));
  Workspace ws(nullptr, 0);
  ws.take<int32_t>(2 * n_chunks);
  ws.take<float>((size_t)2 * n_chunks * c);
  if (with_arg) ws.take<int32_t>((size_t)2 * n_chunks * c);
  *bytes = ws.used;
  return FSFB_OK;
}

int fsfb_segment_reduce(const float* feat, int64_t n, int c, int64_t feat_stride,
                        const int32_t* perm, const int32_t* seg, const int32_t* offsets, int64_t m,
                        int mode, float* out, int64_t* argmax, void* workspace,
                        size_t workspace_bytes, void* stream) {
  using namespace fsfb;
  FSFB_CHECK_ARG(n >= 0 && n < (1ll << 31) && m >= 0 && m < (1ll << 31) && c >= 1 && feat_stride >= c,
                 "segment_reduce: bad n=%lld m=%lld c=%d stride=%lld", (long long)n, (long long)m, c,
                 (long long)feat_stride);
  const int dense = (mode & FSFB_REDUCE_DENSE) ? 1 : 0;
  mode &= ~FSFB_REDUCE_DENSE;
  FSFB_CHECK_ARG(mode == FSFB_REDUCE_SUM || mode == FSFB_REDUCE_MEAN || mode == FSFB_REDUCE_MAX,
                 "segment_reduce: bad mode %d", mode);
  FSFB_CHECK_ARG(argmax == nullptr || mode == FSFB_REDUCE_MAX, "segment_reduce: argmax needs MAX");
  if (m == 0) return FSFB_OK;
  FSFB_CHECK_ARG(offsets && out && (n == 0 || (feat && seg)), "segment_reduce: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n_chunks = std::max<int64_t>(1, ceil_div(n, kSrChunk));
  Workspace ws(workspace, workspace_bytes);
  int32_t* part_seg = ws.take<int32_t>(2 * n_chunks);
  float* part_val = ws.take<float>((size_t)2 * n_chunks * c);
  int32_t* part_arg = argmax ? ws.take<int32_t>((size_t)2 * n_chunks * c) : nullptr;
  if (!ws.ok()) {
    set_error("segment_reduce: workspace too small (%zu given, %zu needed)", workspace_bytes, ws.used);
    return FSFB_ERR_CAPACITY;
  }
  const bool vec4 = (c % 4 == 0) && (feat_stride % 4 == 0) && ((uintptr_t)feat % 16 == 0) &&
                    ((uintptr_t)out % 16 == 0);
  long long* argout = (long long*)argmax;
  int rc;
  if (n <= (int64_t)kSrSmallAvg * m) {  // short segments on average: warp-per-segment kernel, single pass
#define SRS_DISPATCH(VEC, K) rc = launch_segreduce_small<VEC, K>(feat, feat_stride, c, perm, offsets, (int)m, mode, n, out, argout, st)
    if (vec4 && c == 132 && !argout) {  // 131-channel point features in padded rows: dedicated layout
      static int resident132[2] = {0, 0};
      const bool is_max = mode == FSFB_REDUCE_MAX;
      auto kern = is_max ? k_segreduce_small_132<true> : k_segreduce_small_132<false>;
      int& resident = resident132[is_max ? 1 : 0];
      if (resident == 0) FSFB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, kern, kSrWarps * 32, 0));
      const int grid = (int)std::min<int64_t>(ceil_div(ceil_div(m, 2), kSrWarps), (int64_t)kNumSMs * std::max(1, resident));
      FSFB_LAUNCH(kern, grid, kSrWarps * 32, 0, st, feat, feat_stride, perm, offsets, (int)m, mode == FSFB_REDUCE_MEAN ? 1 : 0, out);
      rc = FSFB_OK;
    } else if (vec4) {
      const int g = (int)ceil_div(c, 128);
      if (g <= 1) SRS_DISPATCH(4, 1);
      else if (g <= 2) SRS_DISPATCH(4, 2);
      else SRS_DISPATCH(4, 4);
    } else {
      const int g = (int)ceil_div(c, 32);
      if (g <= 1) SRS_DISPATCH(1, 1);
      else if (g <= 2) SRS_DISPATCH(1, 2);
      else if (g <= 3) SRS_DISPATCH(1, 3);
      else if (g <= 4) SRS_DISPATCH(1, 4);
      else if (g <= 5) SRS_DISPATCH(1, 5);
      else if (g <= 6) SRS_DISPATCH(1, 6);
      else SRS_DISPATCH(1, 8);
    }
#undef SRS_DISPATCH
    return rc;
  }
#define SR_DISPATCH(VEC, K)                                                                     \
  rc = launch_segreduce<VEC, K>(feat, feat_stride, c, perm, seg, offsets, (int)m, mode, out,     \
                                argout, part_seg, part_val, part_arg, n_chunks, st)
  if (vec4) {
    const int lanes_groups = (int)ceil_div(c, 128);
    if (lanes_groups <= 1) SR_DISPATCH(4, 1);
    else if (lanes_groups <= 2) SR_DISPATCH(4, 2);
    else SR_DISPATCH(4, 4);
  } else {
    const int lanes_groups = (int)ceil_div(c, 32);
    if (lanes_groups <= 1) SR_DISPATCH(1, 1);
    else if (lanes_groups <= 2) SR_DISPATCH(1, 2);
    else if (lanes_groups <= 3) SR_DISPATCH(1, 3);
    else if (lanes_groups <= 4) SR_DISPATCH(1, 4);
    else if (lanes_groups <= 5) SR_DISPATCH(1, 5);
    else if (lanes_groups <= 6) SR_DISPATCH(1, 6);
    else SR_DISPATCH(1, 8);
  }
#undef SR_DISPATCH
  if (rc != FSFB_OK) return rc;
  const int64_t cblocks = ceil_div(c, 32);
  const int64_t items = dense ? n_chunks * cblocks
                              : std::max<int64_t>(n_chunks * cblocks, std::min<int64_t>(m, (int64_t)kNumSMs * 64));
  const int grid = (int)ceil_div(items, kSrWarps);
  const int mean = mode == FSFB_REDUCE_MEAN;
  auto fix = (mode == FSFB_REDUCE_MAX)
                 ? (argmax ? k_segreduce_fixup<true, true> : k_segreduce_fixup<true, false>)
                 : k_segreduce_fixup<false, false>;
  FSFB_LAUNCH(fix, grid, kSrWarps * 32, 0, st, c, offsets, (int)m, mean, dense, n, out, argout, part_seg,
              part_val, part_arg, n_chunks);
  return FSFB_OK;
}

int fsfb_gather_rows(const float* src, int64_t m, int c, int64_t src_stride, const void* idx, int idx_i64, int64_t n,
                     float fill, float* out, int64_t out_stride, void* stream) {
  using namespace fsfb;
  FSFB_CHECK_ARG(n >= 0 && m >= 0 && c >= 1 && out_stride >= c && src_stride >= c, "gather_rows: bad argument");
  if (n == 0) return FSFB_OK;
  FSFB_CHECK_ARG(idx && out && (m == 0 || src), "gather_rows: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec4 = (c % 4 == 0) && (out_stride % 4 == 0) && (src_stride % 4 == 0) && ((uintptr_t)src % 16 == 0) &&
                    ((uintptr_t)out % 16 == 0);
  const int64_t chunks = n * (vec4 ? c / 4 : c);
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(chunks, 256 * 4), (int64_t)kNumSMs * 16));
  if (idx_i64) {
    auto kern = vec4 ? k_gather_rows<long long, 4> : k_gather_rows<long long, 1>;
    FSFB_LAUNCH(kern, grid, 256, 0, st, src, m, c, src_stride, (const long long*)idx, n, fill, out, out_stride);
  } else {
    auto kern = vec4 ? k_gather_rows<int, 4> : k_gather_rows<int, 1>;
    FSFB_LAUNCH(kern, grid, 256, 0, st, src, m, c, src_stride, (const int*)idx, n, fill, out, out_stride);
  }
  return FSFB_OK;
}

}  // extern "C"
