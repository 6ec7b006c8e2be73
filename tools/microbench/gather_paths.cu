// gather_paths.cu — how fast can ONE SM pull scattered feature rows out of L2?  (round-2 prerequisite)
//
// The persistent gather-GEMM gathers 128 rows per stage (row = 128 / 256 / 512 bytes of an L2-resident [M, C] fp32 matrix)
// with thread = row 256-bit loads.  This standalone benchmark measures the alternatives on the same access pattern:
//   0  thread = row, 4 x LDG.256 per 128 bytes        (what the kernel does today)
//   1  warp  = row, one coalesced LDG.128 per lane    (needs a transpose afterwards)
//   2  cp.async 16 B (LDGSTS) thread = row → shared
//   3  cp.async.bulk row copies → shared, completion on an mbarrier (TMA unit, no LSU)
// Every variant runs 148 x CTAS_PER_SM persistent CTAs of 512 threads over the same index list and reports bytes / clk / SM and
// aggregate GB/s; indices are either random (worst case) or locally sorted (neighbour rows of a rulebook are clustered).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/microbench/gather_paths tools/microbench/gather_paths.cu
//   timeout 120 tools/microbench/gather_paths
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int kThreads = 512;
constexpr int kRowsPerStage = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- 0: thread = row, 256-bit loads ---------------------------------------------------------------------------------
template <int ROW_BYTES>
__global__ void __launch_bounds__(kThreads) k_thread_row(const float* __restrict__ a, const int* __restrict__ idx, int64_t n_stages,
                                                         float* __restrict__ sink, unsigned long long* clk) {
  const int tid = threadIdx.x;
  float acc = 0.f;
  const long long t0 = clock64();
  // 512 threads = 4 groups of 128 rows: group g takes stages g, g+4, ... of this CTA's share
  for (int64_t s = (int64_t)blockIdx.x * 4 + (tid >> 7); s < n_stages; s += (int64_t)gridDim.x * 4) {
    const int row = idx[s * kRowsPerStage + (tid & 127)];
    const float* g = a + (int64_t)row * (ROW_BYTES / 4);
#pragma unroll
    for (int j = 0; j < ROW_BYTES / 32; ++j) {
      float v0, v1, v2, v3, v4, v5, v6, v7;
      asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=f"(v0), "=f"(v1), "=f"(v2), "=f"(v3), "=f"(v4), "=f"(v5), "=f"(v6), "=f"(v7)
                   : "l"(g + 8 * j));
      acc += v0 + v1 + v2 + v3 + v4 + v5 + v6 + v7;
    }
  }
  const long long t1 = clock64();
  if (acc == 12345.678f) sink[0] = acc;
  if (tid == 0) clk[blockIdx.x] = (unsigned long long)(t1 - t0);
}

// ---- 1: warp = row, coalesced 128-bit loads -----------------------------------------------------------------------------
template <int ROW_BYTES>
__global__ void __launch_bounds__(kThreads) k_warp_row(const float* __restrict__ a, const int* __restrict__ idx, int64_t n_stages,
                                                       float* __restrict__ sink, unsigned long long* clk) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float acc = 0.f;
  const long long t0 = clock64();
  constexpr int kLanesPerRow = ROW_BYTES / 16;           // 8 / 16 / 32 lanes cover one row
  constexpr int kRowsPerInstr = 32 / kLanesPerRow;       // rows one warp-wide load covers
  for (int64_t s = blockIdx.x; s < n_stages; s += gridDim.x) {
    // 16 warps x 8 rows each
    for (int r = 0; r < 8; r += kRowsPerInstr) {
      const int row = idx[s * kRowsPerStage + warp * 8 + r + lane / kLanesPerRow];
      const float4 v = __ldg(reinterpret_cast<const float4*>(a + (int64_t)row * (ROW_BYTES / 4)) + (lane % kLanesPerRow));
      acc += v.x + v.y + v.z + v.w;
    }
  }
  const long long t1 = clock64();
  if (acc == 12345.678f) sink[0] = acc;
  if (tid == 0) clk[blockIdx.x] = (unsigned long long)(t1 - t0);
}

// ---- 2: cp.async 16 B, thread = row → shared -------------------------------------------------------------------------------
template <int ROW_BYTES>
__global__ void __launch_bounds__(kThreads) k_ldgsts(const float* __restrict__ a, const int* __restrict__ idx, int64_t n_stages,
                                                     float* __restrict__ sink, unsigned long long* clk) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x;
  const uint32_t dst = smem_u32(smem) + (uint32_t)tid * (128 + 16);           // padded 128-byte landing zone per thread (reused)
  const long long t0 = clock64();
  for (int64_t s = (int64_t)blockIdx.x * 4 + (tid >> 7); s < n_stages; s += (int64_t)gridDim.x * 4) {
    const int row = idx[s * kRowsPerStage + (tid & 127)];
    const char* g = reinterpret_cast<const char*>(a + (int64_t)row * (ROW_BYTES / 4));
#pragma unroll
    for (int j = 0; j < ROW_BYTES / 16; ++j)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16 * (j & 7)), "l"(g + 16 * j) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 1;" ::: "memory");                      // one stage in flight behind the current one
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  const long long t1 = clock64();
  if (smem[tid] == 123 && clk == nullptr) sink[0] = 1.f;
  if (tid == 0) clk[blockIdx.x] = (unsigned long long)(t1 - t0);
}

// ---- 3: cp.async.bulk row copies, completion on an mbarrier ---------------------------------------------------------------
// 4 stage slots of 128 rows; warp w < 4 issues the copies of slot w (lane = 4 rows each), everybody else idles:
// the TMA unit does the work, which is the point.
template <int ROW_BYTES>
__global__ void __launch_bounds__(kThreads) k_bulk(const float* __restrict__ a, const int* __restrict__ idx, int64_t n_stages,
                                                   float* __restrict__ sink, unsigned long long* clk) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bar[4];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[i])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const long long t0 = clock64();
  if (warp < 4) {
    const uint32_t b = smem_u32(&bar[warp]);
    const uint32_t slot = smem_u32(smem) + (uint32_t)warp * 32768u;   // 32 KB landing zone per slot (wider rows wrap: data unused)
    uint32_t ph = 0;
    for (int64_t s = (int64_t)blockIdx.x * 4 + warp; s < n_stages; s += (int64_t)gridDim.x * 4) {
      if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(kRowsPerStage * ROW_BYTES) : "memory");
      __syncwarp();
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int rl = 4 * lane + r;
        const int row = idx[s * kRowsPerStage + rl];
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(slot + (uint32_t)(rl * ROW_BYTES) % 32768u),
                     "l"(a + (int64_t)row * (ROW_BYTES / 4)), "r"(ROW_BYTES), "r"(b)
                     : "memory");
      }
      // wait for this slot before reusing it (4 slots = 4 warps: the other three keep the unit busy meanwhile)
      asm volatile(
          "{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(b),
          "r"(ph)
          : "memory");
      ph ^= 1u;
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (smem[tid] == 123 && clk == nullptr) sink[0] = 1.f;
  if (tid == 0) clk[blockIdx.x] = (unsigned long long)(t1 - t0);
}

template <typename K>
static void run(const char* name, K kernel, int row_bytes, const float* a, const int* idx, int64_t n_stages, size_t smem, int grid,
                float* sink, unsigned long long* clk) {
  if (smem) CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  kernel<<<grid, kThreads, smem>>>(a, idx, n_stages, sink, clk);   // warm-up (pulls the matrix into L2)
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  kernel<<<grid, kThreads, smem>>>(a, idx, n_stages, sink, clk);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  std::vector<unsigned long long> h(grid);
  CK(cudaMemcpy(h.data(), clk, grid * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  unsigned long long mx = 0;
  for (auto c : h) mx = std::max(mx, c);
  const double bytes = (double)n_stages * kRowsPerStage * row_bytes;
  printf("  %-28s %4d B rows: %8.1f GB/s  %6.1f B/clk/SM  (%.3f ms, %llu clk)\n", name, row_bytes, bytes / ms * 1e-6,
         bytes / 148.0 / (double)mx, ms, mx);
}

template <int ROW_BYTES>
static void run_all(const float* a, const int* idx, int64_t n_stages, float* sink, unsigned long long* clk) {
  const int grid = 148;
  run("thread=row LDG.256", k_thread_row<ROW_BYTES>, ROW_BYTES, a, idx, n_stages, 0, grid, sink, clk);
  run("warp=row LDG.128", k_warp_row<ROW_BYTES>, ROW_BYTES, a, idx, n_stages, 0, grid, sink, clk);
  run("thread=row cp.async 16B", k_ldgsts<ROW_BYTES>, ROW_BYTES, a, idx, n_stages, (size_t)kThreads * (128 + 16), grid, sink, clk);
  run("cp.async.bulk rows", k_bulk<ROW_BYTES>, ROW_BYTES, a, idx, n_stages, (size_t)4 * 32768, grid, sink, clk);
}

int main() {
  const int64_t M = 100000;           // voxels (L2-resident at every width tested: <= 51 MB)
  const int64_t n_stages = 148 * 64;  // 1.2 M gathered rows
  float* a;
  int* idx;
  float* sink;
  unsigned long long* clk;
  CK(cudaMalloc(&a, M * 512));
  CK(cudaMemset(a, 0, M * 512));
  CK(cudaMalloc(&idx, n_stages * kRowsPerStage * sizeof(int)));
  CK(cudaMalloc(&sink, 16));
  CK(cudaMalloc(&clk, 1024 * sizeof(unsigned long long)));
  std::vector<int> h(n_stages * kRowsPerStage);
  for (int pattern = 0; pattern < 2; ++pattern) {
    uint64_t st = 88172645463325252ull;
    for (auto& v : h) {
      st ^= st << 13; st ^= st >> 7; st ^= st << 17;
      v = (int)(st % (uint64_t)M);
    }
    if (pattern == 1)  // clustered: each stage's 128 rows come from a window of 4096 rows, sorted
      for (int64_t s = 0; s < n_stages; ++s) {
        const int base = h[s * kRowsPerStage] % (int)(M - 4096);
        for (int r = 0; r < kRowsPerStage; ++r) h[s * kRowsPerStage + r] = base + h[s * kRowsPerStage + r] % 4096;
        std::sort(h.begin() + s * kRowsPerStage, h.begin() + (s + 1) * kRowsPerStage);
      }
    CK(cudaMemcpy(idx, h.data(), h.size() * sizeof(int), cudaMemcpyHostToDevice));
    printf("%s indices, %lld rows of a %lld-row matrix, 148 CTAs x 512 threads\n", pattern ? "clustered" : "random",
           (long long)(n_stages * kRowsPerStage), (long long)M);
    run_all<128>(a, idx, n_stages, sink, clk);
    run_all<256>(a, idx, n_stages, sink, clk);
    run_all<512>(a, idx, n_stages, sink, clk);
  }
  return 0;
}
