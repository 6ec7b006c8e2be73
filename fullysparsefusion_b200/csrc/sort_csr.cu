// sort_csr.cu — stable LSD radix sort of (segment id, row) pairs and the segment CSR built
// from it; in-group indices on top of the CSR.
//
// The CSR (offsets / perm / seg) is this framework's "scatter rulebook": it is built once
// per ranking and reused by every segmented reduction over the same ids (SIR runs six
// scatter_max over one id set, models/backbones/sir.py:67-81; pre_voxelize five means,
// models/detectors/single_stage_fsd.py:597-601).  The sort is stable, so inside a segment
// perm is ascending in source row: reductions are deterministic and the in-group index
// equals the reference's slow oracle (models/middle_encoders/sst_input_layer.py:200-208).
//
// Radix passes use up to 11 bits (2048 bins, 64 KB of per-warp counters in shared memory),
// so <= 2048 segments sort in one pass and <= 4M segments in two.
#include <cstdlib>

#include "common.cuh"

namespace fsfb {

constexpr int kRsThreads = 256;
constexpr int kRsWarps = kRsThreads / 32;
constexpr int kRsItems = 16;                       // items per thread
constexpr int kRsTile = kRsThreads * kRsItems;     // 4096 items per CTA
constexpr int kRsWarpSpan = 32 * kRsItems;         // 512 contiguous items per warp
constexpr int kRsMaxBits = 11;

// keys: segment id per row; id < 0 (dropped row) sorts last as key m.
template <typename IdxT>
__device__ __forceinline__ uint32_t load_key(const IdxT* __restrict__ index, int64_t i, uint32_t m) {
  long long v = (long long)index[i];
  return (v < 0 || v >= (long long)m) ? m : (uint32_t)v;
}

// ---- pass kernel 1: per-tile digit histogram ------------------------------------------
// hist layout: [bin][tile] so that one flat exclusive scan gives stable global offsets.
template <typename IdxT, bool kFirst>
__global__ void __launch_bounds__(kRsThreads)
    k_rs_hist(const IdxT* __restrict__ index, const uint32_t* __restrict__ keys_in, int64_t n,
              uint32_t m, int shift, int rbits, uint32_t* __restrict__ hist, int ntiles) {
  extern __shared__ uint32_t s_cnt[];
  const int nbins = 1 << rbits;
  for (int b = threadIdx.x; b < nbins; b += kRsThreads) s_cnt[b] = 0;
  __syncthreads();
  int64_t base = (int64_t)blockIdx.x * kRsTile;
#pragma unroll 4
  for (int k = 0; k < kRsItems; ++k) {
    int64_t i = base + (int64_t)k * kRsThreads + threadIdx.x;
    if (i < n) {
      uint32_t key = kFirst ? load_key(index, i, m) : keys_in[i];
      atomicAdd(&s_cnt[(key >> shift) & (nbins - 1)], 1u);
    }
  }
  __syncthreads();
  for (int b = threadIdx.x; b < nbins; b += kRsThreads) hist[(size_t)b * ntiles + blockIdx.x] = s_cnt[b];
}

// ---- pass kernel 2: exclusive scan of hist in [bin][tile] order, two levels ---------------------
// (a) one CTA per bin scans that bin's per-tile counts in place and records the bin total;
// (b) one CTA scans the <= 2048 bin totals.  Global offset of (bin, tile) = bin_base[bin] + hist[bin][tile].
__global__ void __launch_bounds__(256) k_rs_scan_bins(uint32_t* __restrict__ hist, int ntiles,
                                                      uint32_t* __restrict__ bin_total) {
  __shared__ uint32_t s_warp[8];
  __shared__ uint32_t s_carry;
  uint32_t* h = hist + (size_t)blockIdx.x * ntiles;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (int base = 0; base < ntiles; base += 256) {
    const int i = base + threadIdx.x;
    const uint32_t v = i < ntiles ? h[i] : 0;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
      if ((int)lane_id() >= o) x += y;
    }
    const int w = threadIdx.x >> 5;
    const uint32_t carry = s_carry;
    if (lane_id() == 31) s_warp[w] = x;
    __syncthreads();
    uint32_t woff = 0, total = 0;
#pragma unroll
    for (int ww = 0; ww < 8; ++ww) {
      const uint32_t t = s_warp[ww];
      if (ww < w) woff += t;
      total += t;
    }
    if (i < ntiles) h[i] = carry + woff + x - v;
    __syncthreads();
    if (threadIdx.x == 0) s_carry = carry + total;
    __syncthreads();
  }
  if (threadIdx.x == 0) bin_total[blockIdx.x] = s_carry;
}

__global__ void __launch_bounds__(1024) k_rs_scan_totals(uint32_t* __restrict__ bin_total, int nbins) {
  __shared__ uint32_t s_warp[32];
  uint32_t v[2];
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int i = threadIdx.x * 2 + k;
    v[k] = i < nbins ? bin_total[i] : 0;
    s += v[k];
  }
  uint32_t x = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
    if ((int)lane_id() >= o) x += y;
  }
  const int w = threadIdx.x >> 5;
  if (lane_id() == 31) s_warp[w] = x;
  __syncthreads();
  if (w == 0) {
    const uint32_t t = s_warp[lane_id()];
    uint32_t u = t;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, u, o);
      if ((int)lane_id() >= o) u += y;
    }
    s_warp[lane_id()] = u - t;
  }
  __syncthreads();
  uint32_t ex = s_warp[w] + x - s;
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int i = threadIdx.x * 2 + k;
    if (i < nbins) bin_total[i] = ex;
    ex += v[k];
  }
}

// ---- pass kernel 3: stable scatter -----------------------------------------------------
template <typename IdxT, bool kFirst>
__global__ void __launch_bounds__(kRsThreads)
    k_rs_scatter(const IdxT* __restrict__ index, const uint32_t* __restrict__ keys_in,
                 const uint32_t* __restrict__ vals_in, uint32_t* __restrict__ keys_out,
                 uint32_t* __restrict__ vals_out, int64_t n, uint32_t m, int shift, int rbits,
                 const uint32_t* __restrict__ hist, const uint32_t* __restrict__ bin_base, int ntiles) {
  extern __shared__ uint32_t s_wcnt[];  // [kRsWarps][nbins]
  const int nbins = 1 << rbits;
  for (int b = threadIdx.x; b < nbins * kRsWarps; b += kRsThreads) s_wcnt[b] = 0;
  __syncthreads();
  const int w = threadIdx.x >> 5, lane = lane_id();
  uint32_t* my = s_wcnt + w * nbins;
  const int64_t wbase = (int64_t)blockIdx.x * kRsTile + (int64_t)w * kRsWarpSpan;
  uint32_t key[kRsItems], val[kRsItems], rnk[kRsItems];
#pragma unroll
  for (int k = 0; k < kRsItems; ++k) {
    int64_t i = wbase + k * 32 + lane;
    bool valid = i < n;
    key[k] = 0;
    val[k] = 0;
    if (valid) {
      key[k] = kFirst ? load_key(index, i, m) : keys_in[i];
      val[k] = kFirst ? (uint32_t)i : vals_in[i];
    }
    // lanes past n get a digit no one shares so they form singleton match groups
    uint32_t digit = valid ? ((key[k] >> shift) & (nbins - 1)) : (0x10000u | lane);
    unsigned peers = __match_any_sync(0xffffffffu, digit);
    int leader = __ffs(peers) - 1;
    uint32_t old = 0;
    if (lane == leader && valid) {
      old = my[digit];
      my[digit] = old + __popc(peers);
    }
    old = __shfl_sync(0xffffffffu, old, leader);
    rnk[k] = old + __popc(peers & lanemask_lt());
    __syncwarp();
  }
  __syncthreads();
  // cross-warp exclusive offsets per digit, plus the tile's global offset for that digit
  for (int b = threadIdx.x; b < nbins; b += kRsThreads) {
    uint32_t run = hist[(size_t)b * ntiles + blockIdx.x] + bin_base[b];
#pragma unroll
    for (int ww = 0; ww < kRsWarps; ++ww) {
      uint32_t c = s_wcnt[ww * nbins + b];
      s_wcnt[ww * nbins + b] = run;
      run += c;
    }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < kRsItems; ++k) {
    int64_t i = wbase + k * 32 + lane;
    if (i < n) {
      uint32_t digit = (key[k] >> shift) & (nbins - 1);
      uint32_t dst = my[digit] + rnk[k];
      keys_out[dst] = key[k];
      vals_out[dst] = val[k];
    }
  }
}

// ---- CSR offsets from sorted segment ids -----------------------------------------------
// seg[j] sorted ascending in [0, m] (m = dropped).  offsets[s] = first j with seg[j] >= s.
__global__ void __launch_bounds__(256)
    k_csr_offsets(const uint32_t* __restrict__ seg, int64_t n, uint32_t m,
                  int32_t* __restrict__ offsets) {
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j <= n;
       j += (int64_t)gridDim.x * blockDim.x) {
    // boundary between position j-1 and j
    long long prev = (j == 0) ? -1 : (long long)seg[j - 1];
    long long cur = (j == n) ? (long long)m : (long long)seg[j];
    if (cur > (long long)m) cur = m;
    for (long long s = prev + 1; s <= cur; ++s) offsets[s] = (int32_t)j;
  }
}

__global__ void __launch_bounds__(256)
    k_ingroup_from_csr(const uint32_t* __restrict__ seg, const uint32_t* __restrict__ perm,
                       const int32_t* __restrict__ offsets, int64_t n, uint32_t m,
                       long long* __restrict__ out) {
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n;
       j += (int64_t)gridDim.x * blockDim.x) {
    uint32_t s = seg[j];
    out[perm[j]] = (s >= m) ? -1ll : (long long)(j - offsets[s]);
  }
}

// Sort key of an output row, so that rows with similar sets of neighbour offsets become adjacent AND tiles come out ordered by
// cost (the persistent gather-GEMM hands out tiles from the end of the order: longest first).
//   koff == 27 (3x3x3 kernels, offsets z-major), mode 1 (default): (any neighbour in the upper plane, any in the lower plane,
//     the 27-bit mask) = 29 bits.  tools/tile_reuse_study.py: a 128-row tile then visits 9.2 offsets on the 160 k-voxel level
//     of the synthetic frame against 11.1 with the popcount-major key (and 3-14 % fewer on the strided and coarser levels) —
//     surfaces split into rows that only have in-plane neighbours and rows that also have vertical ones;
//   mode 0 (FSFB_ROW_KEY=0, the key the round-1 measurements were taken with) and koff <= 26: popcount << koff | mask;
//   koff > 27: the mask alone.
__host__ __device__ inline uint32_t rulebook_key(uint32_t m, int koff, int mode) {
#ifdef __CUDA_ARCH__
  const uint32_t pc = (uint32_t)__popc(m);
#else
  const uint32_t pc = (uint32_t)__builtin_popcount(m);
#endif
  if (koff <= 26) return (pc << koff) | m;
  if (koff == 27) {
    if (mode == 0) return (pc << 26) | (m & 0x1FFFu) | ((m >> 14) << 13);  // centre bit (13) dropped to make room
    const uint32_t up = (m >> 18) != 0u ? 1u : 0u, dn = (m & 0x1FFu) != 0u ? 1u : 0u;
    return (up << 28) | (dn << 27) | m;
  }
  return m;
}
inline int rulebook_key_mode() {
  static const int mode = [] {
    const char* e = getenv("FSFB_ROW_KEY");
    return e ? atoi(e) : 1;
  }();
  return mode;
}
__global__ void __launch_bounds__(256)
    k_rulebook_masks(const int32_t* __restrict__ nbr, int koff, int64_t rows, int32_t* __restrict__ mask, int mode) {
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
    uint32_t m = 0;
    for (int k = 0; k < koff; ++k) m |= (__ldg(nbr + (int64_t)k * rows + r) >= 0 ? 1u : 0u) << k;
    mask[r] = (int32_t)rulebook_key(m, koff, mode);
  }
}

static int bits_for(uint64_t max_value) {
  int b = 1;
  while (b < 32 && (max_value >> b) != 0) ++b;
  return b;
}

// Sort (index → keys in [0,m]) stably; results land in keys_out/vals_out.
// tmp_keys/tmp_vals: ping-pong buffers [n]; hist: [2^rbits * ntiles].
template <typename IdxT>
static int radix_sort_index(const IdxT* index, int64_t n, uint32_t m, uint32_t* keys_out,
                            uint32_t* vals_out, uint32_t* tmp_keys, uint32_t* tmp_vals,
                            uint32_t* hist, cudaStream_t st) {
  const int bits = bits_for(m);
  const int passes = (bits + kRsMaxBits - 1) / kRsMaxBits;
  const int rbits = (bits + passes - 1) / passes;
  const int nbins = 1 << rbits;
  const int ntiles = (int)ceil_div(n, kRsTile);
  const size_t smem_hist = (size_t)nbins * 4, smem_scatter = (size_t)nbins * kRsWarps * 4;
  static bool attr_set = false;
  if (!attr_set) {
    FSFB_CUDA(cudaFuncSetAttribute(k_rs_scatter<int, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    FSFB_CUDA(cudaFuncSetAttribute(k_rs_scatter<int, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    FSFB_CUDA(cudaFuncSetAttribute(k_rs_scatter<long long, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    FSFB_CUDA(cudaFuncSetAttribute(k_rs_scatter<long long, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    attr_set = true;
  }
  // arrange ping-pong so the last pass writes keys_out/vals_out
  uint32_t* kin = nullptr;
  uint32_t* vin = nullptr;
  for (int p = 0; p < passes; ++p) {
    const bool last_to_out = ((passes - 1 - p) % 2) == 0;
    uint32_t* kout = last_to_out ? keys_out : tmp_keys;
    uint32_t* vout = last_to_out ? vals_out : tmp_vals;
    const int shift = p * rbits;
    auto hist_kernel = (p == 0) ? k_rs_hist<IdxT, true> : k_rs_hist<IdxT, false>;
    auto scatter_kernel = (p == 0) ? k_rs_scatter<IdxT, true> : k_rs_scatter<IdxT, false>;
    FSFB_LAUNCH(hist_kernel, ntiles, kRsThreads, smem_hist, st, index, kin, n, m, shift, rbits,
                hist, ntiles);
    uint32_t* bin_base = hist + (size_t)nbins * ntiles;
    FSFB_LAUNCH(k_rs_scan_bins, nbins, 256, 0, st, hist, ntiles, bin_base);
    FSFB_LAUNCH(k_rs_scan_totals, 1, 1024, 0, st, bin_base, nbins);
    FSFB_LAUNCH(scatter_kernel, ntiles, kRsThreads, smem_scatter, st, index, kin, vin, kout, vout,
                n, m, shift, rbits, hist, bin_base, ntiles);
    kin = kout;
    vin = vout;
  }
  return FSFB_OK;
}

static size_t csr_ws_layout(int64_t n, int64_t m, Workspace& ws, uint32_t** tk, uint32_t** tv,
                            uint32_t** hist) {
  const int bits = bits_for((uint64_t)m);
  const int passes = (bits + kRsMaxBits - 1) / kRsMaxBits;
  const int rbits = (bits + passes - 1) / passes;
  const int64_t ntiles = std::max<int64_t>(1, ceil_div(n, kRsTile));
  uint32_t* a = ws.take<uint32_t>(std::max<int64_t>(n, 1));
  uint32_t* b = ws.take<uint32_t>(std::max<int64_t>(n, 1));
  uint32_t* h = ws.take<uint32_t>((size_t)(1 << rbits) * (ntiles + 1));
  if (tk) *tk = a;
  if (tv) *tv = b;
  if (hist) *hist = h;
  return ws.used;
}

static int csr_build_impl(const void* index, int index_i64, int64_t n, int64_t m,
                          int32_t* offsets, uint32_t* perm, uint32_t* seg, void* workspace,
                          size_t workspace_bytes, cudaStream_t st) {
  Workspace ws(workspace, workspace_bytes);
  uint32_t *tk, *tv, *hist;
  csr_ws_layout(n, m, ws, &tk, &tv, &hist);
  if (!ws.ok()) {
    set_error("csr_build: workspace too small (%zu given, %zu needed)", workspace_bytes, ws.used);
    return FSFB_ERR_CAPACITY;
  }
  if (n == 0) {
    FSFB_CUDA(cudaMemsetAsync(offsets, 0, (size_t)(m + 1) * 4, st));
    return FSFB_OK;
  }
  int rc = index_i64 ? radix_sort_index<long long>((const long long*)index, n, (uint32_t)m, seg,
                                                   perm, tk, tv, hist, st)
                     : radix_sort_index<int>((const int*)index, n, (uint32_t)m, seg, perm, tk, tv,
                                             hist, st);
  if (rc != FSFB_OK) return rc;
  int grid = (int)std::min<int64_t>(ceil_div(n + 1, 256), (int64_t)kNumSMs * 8);
  FSFB_LAUNCH(k_csr_offsets, grid, 256, 0, st, seg, n, (uint32_t)m, offsets);
  return FSFB_OK;
}

}  // namespace fsfb

namespace fsfb {
__global__ void __launch_bounds__(256)
    k_permute_rulebook(const int32_t* __restrict__ nbr, int koff, int64_t rows, int64_t stride, const int32_t* __restrict__ order,
                       int32_t* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < stride; i += (int64_t)gridDim.x * 256) {
    const int64_t src = i < rows ? (int64_t)__ldg(order + i) : -1;
    for (int k = 0; k < koff; ++k) out[(int64_t)k * stride + i] = src >= 0 ? __ldg(nbr + (int64_t)k * rows + src) : -1;
  }
}
}  // namespace fsfb

extern "C" {

int fsfb_csr_workspace_bytes(int64_t n, int64_t m, size_t* bytes) {
  using namespace fsfb;
  FSFB_CHECK_ARG(bytes && n >= 0 && m >= 0 && m < (1ll << 31) - 1 && n < (1ll << 31),
                 "csr_workspace_bytes: bad argument");
  Workspace ws(nullptr, 0);
  *bytes = csr_ws_layout(n, m, ws, nullptr, nullptr, nullptr);
  return FSFB_OK;
}

int fsfb_csr_build(const void* index, int index_i64, int64_t n, int64_t m, int32_t* offsets,
                   int32_t* perm, int32_t* seg, void* workspace, size_t workspace_bytes,
                   void* stream) {
  using namespace fsfb;
  FSFB_CHECK_ARG(n >= 0 && m >= 0 && m < (1ll << 31) - 1 && n < (1ll << 31),
                 "csr_build: bad n=%lld m=%lld", (long long)n, (long long)m);
  FSFB_CHECK_ARG(offsets && (n == 0 || (index && perm && seg)), "csr_build: null pointer");
  return csr_build_impl(index, index_i64, n, m, offsets, (uint32_t*)perm, (uint32_t*)seg, workspace,
                        workspace_bytes, (cudaStream_t)stream);
}

int fsfb_rulebook_order_workspace_bytes(int64_t rows, int koff, size_t* bytes) {
  using namespace fsfb;
  FSFB_CHECK_ARG(bytes && rows >= 0 && rows < (1ll << 31) && koff >= 1 && koff <= 30, "rulebook_order_workspace_bytes: bad argument");
  Workspace ws(nullptr, 0);
  ws.take<int32_t>(std::max<int64_t>(rows, 1));   // masks
  ws.take<uint32_t>(std::max<int64_t>(rows, 1));  // sorted keys
  csr_ws_layout(rows, (int64_t)rulebook_key((1u << koff) - 1u, koff, 0), ws, nullptr, nullptr, nullptr);  // mode 0 has the larger maximum
  *bytes = ws.used;
  return FSFB_OK;
}

int fsfb_rulebook_row_order(const int32_t* nbr, int koff, int64_t rows, int32_t* order, void* workspace,
                            size_t workspace_bytes, void* stream) {
  using namespace fsfb;
  FSFB_CHECK_ARG(rows >= 0 && rows < (1ll << 31) && koff >= 1 && koff <= 30, "rulebook_row_order: bad argument");
  if (rows == 0) return FSFB_OK;
  FSFB_CHECK_ARG(nbr && order, "rulebook_row_order: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  Workspace ws(workspace, workspace_bytes);
  int32_t* mask = ws.take<int32_t>(rows);
  uint32_t* keys = ws.take<uint32_t>(rows);
  uint32_t *tk, *tv, *hist;
  csr_ws_layout(rows, (int64_t)rulebook_key((1u << koff) - 1u, koff, 0), ws, &tk, &tv, &hist);
  if (!ws.ok()) {
    set_error("rulebook_row_order: workspace too small (%zu given, %zu needed)", workspace_bytes, ws.used);
    return FSFB_ERR_CAPACITY;
  }
  const int grid = (int)std::min<int64_t>(ceil_div(rows, 256), (int64_t)kNumSMs * 8);
  const int mode = rulebook_key_mode();
  FSFB_LAUNCH(k_rulebook_masks, grid, 256, 0, st, nbr, koff, rows, mask, mode);
  // stable LSD sort of the keys: rows with the same set of neighbour offsets become adjacent, fewest offsets first
  return radix_sort_index<int>(mask, rows, rulebook_key((1u << koff) - 1u, koff, mode), keys, (uint32_t*)order, tk, tv, hist, st);
}

/* out[k][i] = nbr[k][order[i]] for i < rows, -1 for rows <= i < stride = round_up(rows, 128): the neighbour table in the row order,
 * padded to whole tiles (FSFB_NBR_ROW_ORDERED, include/fsf_b200.h). */
int fsfb_permute_rulebook(const int32_t* nbr, int koff, int64_t rows, const int32_t* order, int32_t* out, void* stream) {
  using namespace fsfb;
  FSFB_CHECK_ARG(rows >= 0 && rows < (1ll << 31) && koff >= 1 && koff <= 32, "permute_rulebook: bad argument");
  if (rows == 0) return FSFB_OK;
  FSFB_CHECK_ARG(nbr && order && out, "permute_rulebook: null pointer");
  const int64_t stride = ceil_div(rows, 128) * 128;
  const int grid = (int)std::min<int64_t>(ceil_div(stride, 256), (int64_t)kNumSMs * 8);
  FSFB_LAUNCH(k_permute_rulebook, grid, 256, 0, (cudaStream_t)stream, nbr, koff, rows, stride, order, out);
  return FSFB_OK;
}

int fsfb_ingroup_workspace_bytes(int64_t n, int64_t m, size_t* bytes) {
  using namespace fsfb;
  FSFB_CHECK_ARG(bytes && n >= 0 && m >= 0 && m < (1ll << 31) - 1 && n < (1ll << 31),
                 "ingroup_workspace_bytes: bad argument");
  Workspace ws(nullptr, 0);
  ws.take<int32_t>(m + 1);
  ws.take<uint32_t>(std::max<int64_t>(n, 1));
  ws.take<uint32_t>(std::max<int64_t>(n, 1));
  csr_ws_layout(n, m, ws, nullptr, nullptr, nullptr);
  *bytes = ws.used;
  return FSFB_OK;
}

int fsfb_ingroup_indices(const int64_t* group, int64_t n, int64_t m, int64_t* out, void* workspace,
                         size_t workspace_bytes, void* stream) {
  using namespace fsfb;
  FSFB_CHECK_ARG(n >= 0 && m >= 0 && m < (1ll << 31) - 1 && n < (1ll << 31),
                 "ingroup_indices: bad n=%lld m=%lld", (long long)n, (long long)m);
  if (n == 0) return FSFB_OK;
  FSFB_CHECK_ARG(group && out, "ingroup_indices: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  Workspace ws(workspace, workspace_bytes);
  int32_t* offsets = ws.take<int32_t>(m + 1);
  uint32_t* perm = ws.take<uint32_t>(n);
  uint32_t* seg = ws.take<uint32_t>(n);
  if (!ws.ok()) {
    set_error("ingroup_indices: workspace too small");
    return FSFB_ERR_CAPACITY;
  }
  int rc = csr_build_impl(group, 1, n, m, offsets, perm, seg, (char*)workspace + ws.used,
                          workspace_bytes - ws.used, st);
  if (rc != FSFB_OK) return rc;
  int grid = (int)std::min<int64_t>(ceil_div(n, 256), (int64_t)kNumSMs * 8);
  FSFB_LAUNCH(k_ingroup_from_csr, grid, 256, 0, st, seg, perm, offsets, n, (uint32_t)m,
              (long long*)out);
  return FSFB_OK;
}

}  // extern "C"
