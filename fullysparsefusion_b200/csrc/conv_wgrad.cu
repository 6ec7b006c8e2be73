// conv_wgrad.cu — weight gradient of the gather-GEMM (sparse conv wgrad for the training step).
//
//     dw[k][co][ci] = sum over output rows r with nbr[k][r] >= 0 of  dy[r][co] * a[nbr[k][r]][ci]      (nbr null: Linear, k = 0)
//
// Reference: the backward of spconv's SubMConv3d / SparseConv3d / SparseInverseConv3d (un-vendored; the modules are built by
// SimpleSparseUNet, config FSF_nuScenes_config.py:58-70) and of nn.Linear in build_mlp (ops/sst_ops.py:808-833); training entry
// tools/train.py:244-251.  The restatement is the definition above; tests compare with per-offset torch.matmul in fp64.
//
// fp32 CUDA-core kernel, deterministic (no atomics): the reduction dimension is the PAIR list of an offset, which is sparse in
// the rows (~20 % of the rows have a given neighbour), so a CTA first compacts the valid (row, source) pairs of a 256-row block
// into shared memory and then runs 32-pair chunks: dy rows and gathered a rows staged as [32][64] tiles, 256 threads with a 4 x 4
// register tile each (two 128-bit shared-memory reads per 16 FMAs).  Grid = (row splits, cout tiles x cin tiles, offsets); the
// splits write partial tiles that a second kernel sums in a fixed order.  A tcgen05 version needs MN-major operand tiles (the
// reduction runs over rows, which are the slow dimension of both operands); not built — DESIGN.md section 8.
#include "common.cuh"

namespace fsfb {

constexpr int kWgTile = 64;      // cout tile = cin tile
constexpr int kWgChunk = 32;     // pairs per FMA chunk
constexpr int kWgBlockRows = 256;

__global__ void __launch_bounds__(256)
    k_conv_wgrad(const float* __restrict__ a, int64_t a_rows, int cin, int64_t a_stride, const float* __restrict__ dy, int64_t rows,
                 int cout, int64_t dy_stride, const int32_t* __restrict__ nbr, int ci_tiles, int64_t rows_per_split,
                 float* __restrict__ part) {
  __shared__ __align__(16) float s_dy[kWgChunk][kWgTile + 4];
  __shared__ __align__(16) float s_a[kWgChunk][kWgTile + 4];
  __shared__ int s_row[kWgBlockRows], s_src[kWgBlockRows];
  __shared__ int s_warp[8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int k = blockIdx.z, split = blockIdx.x;
  const int co0 = (blockIdx.y / ci_tiles) * kWgTile, ci0 = (blockIdx.y % ci_tiles) * kWgTile;
  const int ty = tid >> 4, tx = tid & 15;   // 4 x 4 outputs: co = co0 + 4 ty + i, ci = ci0 + 4 tx + j
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int64_t r_beg = (int64_t)split * rows_per_split, r_end = min(rows, r_beg + rows_per_split);
  for (int64_t rb = r_beg; rb < r_end; rb += kWgBlockRows) {
    // ---- compact the valid pairs of this 256-row block (ascending rows: a fixed summation order) ----
    const int64_t r = rb + tid;
    int src = -1;
    if (r < r_end) src = nbr ? __ldg(nbr + (int64_t)k * rows + r) : (int)r;
    const bool ok = src >= 0 && src < a_rows;
    const unsigned bal = __ballot_sync(0xffffffffu, ok);
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();   // (also: the previous block's chunks are done with s_row / s_src)
    int before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      const int c = s_warp[w];
      before += w < warp ? c : 0;
      total += c;
    }
    if (ok) {
      const int pos = before + __popc(bal & ((1u << lane) - 1u));
      s_row[pos] = (int)(r - rb);
      s_src[pos] = src;
    }
    __syncthreads();
    // ---- 32-pair chunks ----
    for (int p0 = 0; p0 < total; p0 += kWgChunk) {
      const int np = min(kWgChunk, total - p0);
      // stage dy[row][co0 .. co0+63] and a[src][ci0 .. ci0+63]: thread = (pair p = tid / 8, 8 floats at 8 (tid % 8))
      {
        const int p = tid >> 3, q = (tid & 7) * 8;
        const bool live = p < np;
        const float* dr = dy + (rb + (live ? s_row[p0 + p] : 0)) * dy_stride + co0 + q;
        const float* ar = a + (int64_t)(live ? s_src[p0 + p] : 0) * a_stride + ci0 + q;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          s_dy[p][q + e] = (live && co0 + q + e < cout) ? __ldg(dr + e) : 0.f;
          s_a[p][q + e] = (live && ci0 + q + e < cin) ? __ldg(ar + e) : 0.f;
        }
      }
      __syncthreads();
#pragma unroll 8
      for (int p = 0; p < kWgChunk; ++p) {
        const float4 d4 = *reinterpret_cast<const float4*>(&s_dy[p][4 * ty]);
        const float4 a4 = *reinterpret_cast<const float4*>(&s_a[p][4 * tx]);
        const float d[4] = {d4.x, d4.y, d4.z, d4.w}, av[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(d[i], av[j], acc[i][j]);
      }
      __syncthreads();
    }
  }
  // partial tile of this split: part[split][k][cout][cin]
  float* out = part + ((int64_t)split * gridDim.z + k) * cout * cin;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int co = co0 + 4 * ty + i;
    if (co >= cout) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int ci = ci0 + 4 * tx + j;
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

static int wgrad_splits(int64_t rows, int koff, int tiles) {
  const int64_t by_rows = std::max<int64_t>(1, rows / 2048);                          // at least eight row blocks per split
  const int64_t want = std::max<int64_t>(1, ceil_div((int64_t)kNumSMs * 4, (int64_t)koff * tiles));
  return (int)std::min<int64_t>(std::min(by_rows, want), 64);
}

}  // namespace fsfb

extern "C" {

int fsfb_conv_wgrad_workspace_bytes(int64_t rows, int koff, int cin, int cout, size_t* bytes) {
  using namespace fsfb;
  FSFB_CHECK_ARG(bytes && rows >= 0 && koff >= 1 && cin >= 1 && cout >= 1, "conv_wgrad_workspace_bytes: bad argument");
  const int tiles = (int)(ceil_div(cout, kWgTile) * ceil_div(cin, kWgTile));
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
  const int ci_tiles = (int)ceil_div(cin, kWgTile);
  const int tiles = (int)ceil_div(cout, kWgTile) * ci_tiles;
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
  FSFB_LAUNCH(k_conv_wgrad, grid, 256, 0, st, a, a_rows, cin, a_stride, dy, rows, cout, dy_stride, nbr, ci_tiles, rows_per_split, part);
  if (splits > 1) {
    const int rgrid = (int)std::min<int64_t>(ceil_div(n, 256), (int64_t)kNumSMs * 8);
    FSFB_LAUNCH(k_conv_wgrad_reduce, rgrid, 256, 0, st, part, n, splits, dw);
  }
  return FSFB_OK;
}

}  // extern "C"
