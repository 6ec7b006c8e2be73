// conv_wgrad.cu — weight gradient of the gather-GEMM (sparse conv wgrad for the training step).
//
//     dw[k][co][ci] = sum over output rows r with nbr[k][r] >= 0 of  dy[r][co] * a[nbr[k][r]][ci]      (nbr null: Linear, k = 0)
//
// Reference: the backward of spconv's SubMConv3d / SparseConv3d / SparseInverseConv3d (un-vendored; the modules are built by
// SimpleSparseUNet, config FSF_nuScenes_config.py:58-70) and of nn.Linear in build_mlp (ops/sst_ops.py:808-833); training entry
// tools/train.py:244-251.  The restatement is the definition above; tests compare with per-offset torch.matmul in fp64.
//
// fp32 CUDA-core kernel, deterministic (no atomics): the reduction dimension is the PAIR list of an offset, which is sparse in
// the rows (~20 % of the rows have a given neighbour), so a CTA first compacts the valid (row, source) pairs of a 1024-row block
// into shared memory and then runs 32-pair chunks: dy rows and gathered a rows staged as [32][128] tiles (the next chunk's rows
// are in flight in registers meanwhile), 256 threads with an 8 x 8 register tile each (four 128-bit shared-memory reads per 64
// FMAs; 64 x 64 tiles / 4 x 4 per thread for narrow layers).  Grid = (row splits, cout tiles x cin tiles, offsets); the
// splits write partial tiles that a second kernel sums in a fixed order.  A tcgen05 version needs MN-major operand tiles (the
// reduction runs over rows, which are the slow dimension of both operands); not built — DESIGN.md section 8.
#include "common.cuh"

namespace fsfb {

constexpr int kWgChunk = 32;        // pairs per FMA chunk
constexpr int kWgBlockRows = 1024;  // rows compacted at a time (four per thread)

// T x T outputs per thread, 16 x 16 threads: tiles of 64 x 64 (T = 4) or 128 x 128 (T = 8) weights.
// VEC: both operands allow 128-bit loads (16-byte aligned bases, strides and widths that are multiples of 4).
template <int T, bool VEC>
__global__ void __launch_bounds__(256, T == 8 ? 2 : 4)
    k_conv_wgrad(const float* __restrict__ a, int64_t a_rows, int cin, int64_t a_stride, const float* __restrict__ dy, int64_t rows,
                 int cout, int64_t dy_stride, const int32_t* __restrict__ nbr, int ci_tiles, int64_t rows_per_split,
                 float* __restrict__ part) {
  constexpr int kTile = 16 * T;
  constexpr int kV = kTile / 32;   // float4 pieces per thread, operand and chunk (32 pairs x kTile floats / 256 threads / 4)
  __shared__ __align__(16) float s_dy[kWgChunk][kTile + 4];
  __shared__ __align__(16) float s_a[kWgChunk][kTile + 4];
  __shared__ int s_row[kWgBlockRows], s_src[kWgBlockRows];
  __shared__ int s_warp[8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int k = blockIdx.z, split = blockIdx.x;
  const int co0 = (blockIdx.y / ci_tiles) * kTile, ci0 = (blockIdx.y % ci_tiles) * kTile;
  // outputs of a thread: co = co0 + 4 ty + (i & 3) + 64 (i >> 2), ci likewise with tx — groups of four 64 apart, so that the
  // sixteen tx lanes of a 128-bit shared-memory read touch 256 contiguous bytes (no bank conflicts)
  const int ty = tid >> 4, tx = tid & 15;
  float acc[T][T];
#pragma unroll
  for (int i = 0; i < T; ++i)
#pragma unroll
    for (int j = 0; j < T; ++j) acc[i][j] = 0.f;
  // staging: thread = (pair p = tid / 8, float4 pieces (tid % 8) + 8 v of the pair's two tile rows)
  const int sp = tid >> 3, sq = tid & 7;
  float4 rd[kV], ra[kV];
  auto fetch = [&](int64_t rb, int p0, int total) {   // chunk [p0, p0 + 32) of the compacted pairs → registers
    const bool live = p0 + sp < total;
    const float* dr = dy + (rb + (live ? s_row[p0 + sp] : 0)) * dy_stride + co0;
    const float* ar = a + (int64_t)(live ? s_src[p0 + sp] : 0) * a_stride + ci0;
#pragma unroll
    for (int v = 0; v < kV; ++v) {
      const int c = 4 * (sq + 8 * v);
      if (VEC) {
        rd[v] = (live && co0 + c < cout) ? __ldg(reinterpret_cast<const float4*>(dr + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
        ra[v] = (live && ci0 + c < cin) ? __ldg(reinterpret_cast<const float4*>(ar + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
      } else {
        rd[v].x = (live && co0 + c < cout) ? __ldg(dr + c) : 0.f;
        rd[v].y = (live && co0 + c + 1 < cout) ? __ldg(dr + c + 1) : 0.f;
        rd[v].z = (live && co0 + c + 2 < cout) ? __ldg(dr + c + 2) : 0.f;
        rd[v].w = (live && co0 + c + 3 < cout) ? __ldg(dr + c + 3) : 0.f;
        ra[v].x = (live && ci0 + c < cin) ? __ldg(ar + c) : 0.f;
        ra[v].y = (live && ci0 + c + 1 < cin) ? __ldg(ar + c + 1) : 0.f;
        ra[v].z = (live && ci0 + c + 2 < cin) ? __ldg(ar + c + 2) : 0.f;
        ra[v].w = (live && ci0 + c + 3 < cin) ? __ldg(ar + c + 3) : 0.f;
      }
    }
  };
  const int64_t r_beg = (int64_t)split * rows_per_split, r_end = min(rows, r_beg + rows_per_split);
  for (int64_t rb = r_beg; rb < r_end; rb += kWgBlockRows) {
    // ---- compact the valid pairs of this row block (ascending rows: a fixed summation order) ----
    int total = 0;
    __syncthreads();   // the previous block's chunks are done with s_row / s_src
#pragma unroll
    for (int pass = 0; pass < kWgBlockRows / 256; ++pass) {
      const int64_t r = rb + pass * 256 + tid;
      int src = -1;
      if (r < r_end) src = nbr ? __ldg(nbr + (int64_t)k * rows + r) : (int)r;
      const bool ok = src >= 0 && src < a_rows;
      const unsigned bal = __ballot_sync(0xffffffffu, ok);
      if (lane == 0) s_warp[warp] = __popc(bal);
      __syncthreads();
      int before = 0, here = 0;
#pragma unroll
      for (int w = 0; w < 8; ++w) {
        const int c = s_warp[w];
        before += w < warp ? c : 0;
        here += c;
      }
      if (ok) {
        const int pos = total + before + __popc(bal & ((1u << lane) - 1u));
        s_row[pos] = pass * 256 + tid;
        s_src[pos] = src;
      }
      total += here;
      __syncthreads();
    }
    // ---- 32-pair chunks; the next chunk's rows are in flight while this one multiplies ----
    if (total > 0) fetch(rb, 0, total);
    for (int p0 = 0; p0 < total; p0 += kWgChunk) {
#pragma unroll
      for (int v = 0; v < kV; ++v) {
        const int c = 4 * (sq + 8 * v);
        *reinterpret_cast<float4*>(&s_dy[sp][c]) = rd[v];
        *reinterpret_cast<float4*>(&s_a[sp][c]) = ra[v];
      }
      __syncthreads();
      if (p0 + kWgChunk < total) fetch(rb, p0 + kWgChunk, total);
#pragma unroll 4
      for (int p = 0; p < kWgChunk; ++p) {
        float d[T], av[T];
#pragma unroll
        for (int i = 0; i < T; i += 4) {
          const float4 d4 = *reinterpret_cast<const float4*>(&s_dy[p][4 * ty + 16 * i]);
          const float4 a4 = *reinterpret_cast<const float4*>(&s_a[p][4 * tx + 16 * i]);
          d[i] = d4.x; d[i + 1] = d4.y; d[i + 2] = d4.z; d[i + 3] = d4.w;
          av[i] = a4.x; av[i + 1] = a4.y; av[i + 2] = a4.z; av[i + 3] = a4.w;
        }
#pragma unroll
        for (int i = 0; i < T; ++i)
#pragma unroll
          for (int j = 0; j < T; ++j) acc[i][j] = fmaf(d[i], av[j], acc[i][j]);
      }
      __syncthreads();
    }
  }
  // partial tile of this split: part[split][k][cout][cin]
  float* out = part + ((int64_t)split * gridDim.z + k) * cout * cin;
#pragma unroll
  for (int i = 0; i < T; ++i) {
    const int co = co0 + 4 * ty + (i & 3) + 64 * (i >> 2);
    if (co >= cout) continue;
#pragma unroll
    for (int j = 0; j < T; ++j) {
      const int ci = ci0 + 4 * tx + (j & 3) + 64 * (j >> 2);
      if (ci < cin) out[(int64_t)co * cin + ci] = acc[i][j];
    }
  }
}

// The same tiles with the operand chunks copied global -> shared memory by 16-byte cp.async into a double buffer (no staging
// registers: the 8 x 8 register tile plus a prefetched chunk did not fit 128 registers).  Needs 128-bit aligned operands.
template <int T>
__global__ void __launch_bounds__(256, 2)
    k_conv_wgrad_async(const float* __restrict__ a, int64_t a_rows, int cin, int64_t a_stride, const float* __restrict__ dy, int64_t rows,
                       int cout, int64_t dy_stride, const int32_t* __restrict__ nbr, int ci_tiles, int64_t rows_per_split,
                       float* __restrict__ part) {
  constexpr int kTile = 16 * T;
  constexpr int kV = kTile / 32;
  constexpr int kLd = kTile + 4;                       // floats per staged row
  constexpr int kBuf = 2 * kWgChunk * kLd;             // floats per buffer: dy rows, then a rows
  extern __shared__ __align__(16) float s_dyn[];       // [2][kBuf] | s_row[1024] | s_src[1024]
  int* s_row = reinterpret_cast<int*>(s_dyn + 2 * kBuf);
  int* s_src = s_row + kWgBlockRows;
  __shared__ int s_warp[8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int k = blockIdx.z, split = blockIdx.x;
  const int co0 = (blockIdx.y / ci_tiles) * kTile, ci0 = (blockIdx.y % ci_tiles) * kTile;
  const int ty = tid >> 4, tx = tid & 15;
  float acc[T][T];
#pragma unroll
  for (int i = 0; i < T; ++i)
#pragma unroll
    for (int j = 0; j < T; ++j) acc[i][j] = 0.f;
  const int sp = tid >> 3, sq = tid & 7;
  const uint32_t s_base = (uint32_t)__cvta_generic_to_shared(s_dyn);
  auto issue = [&](int64_t rb, int p0, int total, int buf) {   // chunk [p0, p0 + 32) → buffer `buf`, one commit group
    const bool live = p0 + sp < total;
    const float* dr = dy + (rb + (live ? s_row[p0 + sp] : 0)) * dy_stride + co0;
    const float* ar = a + (int64_t)(live ? s_src[p0 + sp] : 0) * a_stride + ci0;
    const uint32_t d_dst = s_base + (uint32_t)(buf * kBuf + sp * kLd) * 4u;
    const uint32_t a_dst = d_dst + (uint32_t)(kWgChunk * kLd) * 4u;
#pragma unroll
    for (int v = 0; v < kV; ++v) {
      const int c = 4 * (sq + 8 * v);
      const bool okd = live && co0 + c < cout, oka = live && ci0 + c < cin;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d_dst + (uint32_t)c * 4u), "l"(okd ? dr + c : dy), "r"(okd ? 16 : 0) : "memory");
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(a_dst + (uint32_t)c * 4u), "l"(oka ? ar + c : a), "r"(oka ? 16 : 0) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  const int64_t r_beg = (int64_t)split * rows_per_split, r_end = min(rows, r_beg + rows_per_split);
  for (int64_t rb = r_beg; rb < r_end; rb += kWgBlockRows) {
    int total = 0;
    __syncthreads();   // the previous block's chunks are done with s_row / s_src and both buffers
#pragma unroll
    for (int pass = 0; pass < kWgBlockRows / 256; ++pass) {
      const int64_t r = rb + pass * 256 + tid;
      int src = -1;
      if (r < r_end) src = nbr ? __ldg(nbr + (int64_t)k * rows + r) : (int)r;
      const bool ok = src >= 0 && src < a_rows;
      const unsigned bal = __ballot_sync(0xffffffffu, ok);
      if (lane == 0) s_warp[warp] = __popc(bal);
      __syncthreads();
      int before = 0, here = 0;
#pragma unroll
      for (int w = 0; w < 8; ++w) {
        const int c = s_warp[w];
        before += w < warp ? c : 0;
        here += c;
      }
      if (ok) {
        const int pos = total + before + __popc(bal & ((1u << lane) - 1u));
        s_row[pos] = pass * 256 + tid;
        s_src[pos] = src;
      }
      total += here;
      __syncthreads();
    }
    if (total > 0) issue(rb, 0, total, 0);
    int buf = 0;
    for (int p0 = 0; p0 < total; p0 += kWgChunk, buf ^= 1) {
      if (p0 + kWgChunk < total) {
        issue(rb, p0 + kWgChunk, total, buf ^ 1);
        asm volatile("cp.async.wait_group 1;" ::: "memory");
      } else {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
      }
      __syncthreads();   // every thread's copies of this chunk have landed
      const float* bd = s_dyn + buf * kBuf;
      const float* ba = bd + kWgChunk * kLd;
#pragma unroll 4
      for (int p = 0; p < kWgChunk; ++p) {
        float d[T], av[T];
#pragma unroll
        for (int i = 0; i < T; i += 4) {
          const float4 d4 = *reinterpret_cast<const float4*>(bd + p * kLd + 4 * ty + 16 * i);
          const float4 a4 = *reinterpret_cast<const float4*>(ba + p * kLd + 4 * tx + 16 * i);
          d[i] = d4.x; d[i + 1] = d4.y; d[i + 2] = d4.z; d[i + 3] = d4.w;
          av[i] = a4.x; av[i + 1] = a4.y; av[i + 2] = a4.z; av[i + 3] = a4.w;
        }
#pragma unroll
        for (int i = 0; i < T; ++i)
#pragma unroll
          for (int j = 0; j < T; ++j) acc[i][j] = fmaf(d[i], av[j], acc[i][j]);
      }
      __syncthreads();   // the buffer may be overwritten by the copies of the chunk after next
    }
  }
  float* out = part + ((int64_t)split * gridDim.z + k) * cout * cin;
#pragma unroll
  for (int i = 0; i < T; ++i) {
    const int co = co0 + 4 * ty + (i & 3) + 64 * (i >> 2);
    if (co >= cout) continue;
#pragma unroll
    for (int j = 0; j < T; ++j) {
      const int ci = ci0 + 4 * tx + (j & 3) + 64 * (j >> 2);
      if (ci < cin) out[(int64_t)co * cin + ci] = acc[i][j];
    }
  }
}

// dw[e] = sum over the splits, in split order
__global__ void __launch_bounds__(256) k_conv_wgrad_reduce(const float* __restrict__ part, int64_t n, int splits, float* __restrict__ dw) {
  for (int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x; e < n; e += (int64_t)gridDim.x * 256) {
    float s = 0.f;
    for (int sp = 0; sp < splits; ++sp) s += part[(int64_t)sp * n + e];
    dw[e] = s;
  }
}

static int wgrad_tile(int cin, int cout) { return (cin >= 96 && cout >= 96) ? 128 : 64; }
static int wgrad_splits(int64_t rows, int koff, int tiles) {
  const int64_t by_rows = std::max<int64_t>(1, rows / (2 * kWgBlockRows));            // at least two row blocks per split
  const int64_t want = std::max<int64_t>(1, ceil_div((int64_t)kNumSMs * 4, (int64_t)koff * tiles));
  return (int)std::min<int64_t>(std::min(by_rows, want), 64);
}

}  // namespace fsfb

extern "C" {

int fsfb_conv_wgrad_workspace_bytes(int64_t rows, int koff, int cin, int cout, size_t* bytes) {
  using namespace fsfb;
  FSFB_CHECK_ARG(bytes && rows >= 0 && koff >= 1 && cin >= 1 && cout >= 1, "conv_wgrad_workspace_bytes: bad argument");
  const int tile = wgrad_tile(cin, cout);
  const int tiles = (int)(ceil_div(cout, tile) * ceil_div(cin, tile));
  const int splits = wgrad_splits(rows, koff, tiles);
  Workspace ws(nullptr, 0);
  if (splits > 1) ws.take<float>((size_t)splits * koff * cout * cin);
  *bytes = ws.used;
  return FSFB_OK;
}

/* dw [koff][cout][cin] f32 (overwritten).  a [a_rows, cin] (row stride a_stride), dy [rows, cout] (row stride dy_stride),
 * nbr [koff][rows] i32 (entries outside [0, a_rows) = no neighbour) or null with koff == 1 (Linear: the pair list is every row). */
int fsfb_conv_wgrad(const float* a, int64_t a_rows, int cin, int64_t a_stride, const float* dy, int64_t rows, int cout,
                    int64_t dy_stride, const int32_t* nbr, int koff, float* dw, void* workspace, size_t workspace_bytes,
                    void* stream) {
  using namespace fsfb;
  FSFB_CHECK_ARG(a_rows >= 0 && a_rows < (1ll << 31) && rows >= 0 && rows < (1ll << 31) && cin >= 1 && cout >= 1 && koff >= 1 &&
                     a_stride >= cin && dy_stride >= cout,
                 "conv_wgrad: bad shape");
  FSFB_CHECK_ARG(nbr || (koff == 1 && a_rows >= rows), "conv_wgrad: several offsets need a neighbour table");
  FSFB_CHECK_ARG(dw, "conv_wgrad: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n = (int64_t)koff * cout * cin;
  if (rows == 0 || a_rows == 0) {
    FSFB_CUDA(cudaMemsetAsync(dw, 0, (size_t)n * sizeof(float), st));
    return FSFB_OK;
  }
  FSFB_CHECK_ARG(a && dy, "conv_wgrad: null pointer");
  const int tile = wgrad_tile(cin, cout);
  const int ci_tiles = (int)ceil_div(cin, tile);
  const int tiles = (int)ceil_div(cout, tile) * ci_tiles;
  const int splits = wgrad_splits(rows, koff, tiles);
  float* part = dw;
  if (splits > 1) {
    Workspace ws(workspace, workspace_bytes);
    part = ws.take<float>((size_t)splits * n);
    if (!ws.ok()) {
      set_error("conv_wgrad: workspace too small (%zu given, %zu needed)", workspace_bytes, ws.used);
      return FSFB_ERR_CAPACITY;
    }
  }
  const int64_t rows_per_split = ceil_div(ceil_div(rows, splits), kWgBlockRows) * kWgBlockRows;
  const dim3 grid((unsigned)splits, (unsigned)tiles, (unsigned)koff);
  const bool vec = ((uintptr_t)a % 16 == 0) && ((uintptr_t)dy % 16 == 0) && a_stride % 4 == 0 && dy_stride % 4 == 0 && cin % 4 == 0 &&
                   cout % 4 == 0;
#define WG_LAUNCH(T, V) \
  FSFB_LAUNCH((k_conv_wgrad<T, V>), grid, 256, 0, st, a, a_rows, cin, a_stride, dy, rows, cout, dy_stride, nbr, ci_tiles, rows_per_split, part)
  if (tile == 128 && vec) {   // the wide tiles of the convolution layers: cp.async double buffer
    constexpr size_t smem = (size_t)(2 * 2 * kWgChunk * (128 + 4)) * 4 + (size_t)2 * kWgBlockRows * 4;
    static bool attr = false;
    if (!attr) {
      FSFB_CUDA(cudaFuncSetAttribute(k_conv_wgrad_async<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr = true;
    }
    FSFB_LAUNCH(k_conv_wgrad_async<8>, grid, 256, smem, st, a, a_rows, cin, a_stride, dy, rows, cout, dy_stride, nbr, ci_tiles,
                rows_per_split, part);
  } else if (tile == 128) {
    WG_LAUNCH(8, false);
  } else {
    if (vec) WG_LAUNCH(4, true); else WG_LAUNCH(4, false);
  }
#undef WG_LAUNCH
  if (splits > 1) {
    const int rgrid = (int)std::min<int64_t>(ceil_div(n, 256), (int64_t)kNumSMs * 8);
    FSFB_LAUNCH(k_conv_wgrad_reduce, rgrid, 256, 0, st, part, n, splits, dw);
  }
  return FSFB_OK;
}

}  // extern "C"
