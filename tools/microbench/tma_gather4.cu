// tma_gather4.cu — can the TMA unit gather the rows of the sparse-convolution operand?  (round-2 evaluation of the north star's
// "TMA-staged shared-memory tiles"; profiles/r2_ncu_summary.md)
//
// cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4: ONE instruction fetches four rows (row coordinates given
// individually) x one box of columns of a 2-d tensor map and lands them, hardware-swizzled, in shared memory, completing on an
// mbarrier.  The operand of k_gather_gemm_ss's pre-split mode is exactly that: 128-byte pieces of 128 gathered rows per stage.
// This benchmark (a) verifies what lands (row order, SWIZZLE_128B placement, zero fill for out-of-range rows = the "no neighbour"
// case) against a host model and (b) measures bytes / clk / SM with one issuing warp per CTA and a 4- or 6-slot ring, next to
// the 16-byte cp.async gather the kernel uses today (4 warps).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/microbench/tma_gather4 tools/microbench/tma_gather4.cu
//   timeout 120 tools/microbench/tma_gather4
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int kRows = 128;          // rows per stage
constexpr int kRowBytes = 128;      // bytes per row piece (one K chunk of the operand)
constexpr int kSlots = 6;
constexpr int kStagesPerCta = 200;   // index lists of a CTA's stages staged in shared memory: 200 x 128 x 4 B = 100 KB next to 6 x 16 KB slots

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_gather4(uint32_t dst, const CUtensorMap* map, uint32_t bar, int col, int r0, int r1, int r2, int r3) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
               ::"r"(dst), "l"(map), "r"(bar), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
}

// ---- gather through the TMA unit: warp 0 issues (lane = 4 rows), warp 1 "consumes" (reads one word per row, frees the slot) ----
__global__ void __launch_bounds__(64) k_tma(const __grid_constant__ CUtensorMap map, const int* __restrict__ idx, int64_t n_stages, int col_pieces,
                                            unsigned long long* clk, uint32_t* check, int slots) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) unsigned long long full[kSlots], empty[kSlots];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kSlots; ++s) { mbar_init(smem_u32(&full[s]), 1); mbar_init(smem_u32(&empty[s]), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // the CTA's row indices live in shared memory, like the neighbour table of the real kernel (no global index load in the loop)
  int* s_idx = reinterpret_cast<int*>(smem + (size_t)kSlots * kRows * kRowBytes);
  const int my_stages = (int)((n_stages - blockIdx.x + gridDim.x - 1) / gridDim.x);
  for (int i = threadIdx.x; i < my_stages * kRows; i += blockDim.x) s_idx[i] = idx[((int64_t)blockIdx.x + (int64_t)(i / kRows) * gridDim.x) * kRows + i % kRows];
  __syncthreads();
  const long long t0 = clock64();
  int s = 0; uint32_t ph = 0;
  uint32_t acc = 0;
  int ls = 0;
  for (int64_t st = blockIdx.x; st < n_stages; st += gridDim.x, ++ls) {
    const uint32_t slot = smem_u32(smem) + (uint32_t)s * kRows * kRowBytes;
    if (warp == 0) {
      if (lane == 0) { mbar_wait(smem_u32(&empty[s]), ph ^ 1u); mbar_expect_tx(smem_u32(&full[s]), kRows * kRowBytes); }
      __syncwarp();
      const int* r = s_idx + ls * kRows + 4 * lane;
      const int col = (int)(st % col_pieces) * kRowBytes;      // which 128-byte piece of the row (the K chunk)
      tma_gather4(slot + (uint32_t)lane * 4 * kRowBytes, &map, smem_u32(&full[s]), col, r[0], r[1], r[2], r[3]);
    } else {
      if (lane == 0) mbar_wait(smem_u32(&full[s]), ph);
      __syncwarp();
      for (int rr = lane; rr < kRows; rr += 32) acc += *reinterpret_cast<const uint32_t*>(smem + (size_t)s * kRows * kRowBytes + rr * kRowBytes);
      if (check && st == blockIdx.x)   // first stage of the CTA: dump the slot for the host-side layout check
        for (int w = lane; w < kRows * kRowBytes / 4; w += 32) check[(size_t)blockIdx.x * (kRows * kRowBytes / 4) + w] = reinterpret_cast<const uint32_t*>(smem + (size_t)s * kRows * kRowBytes)[w];
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&empty[s]));
    }
    if (++s == slots) { s = 0; ph ^= 1u; }
  }
  const long long t1 = clock64();
  if (acc == 0x12345678u && clk == nullptr) printf("x");
  if (threadIdx.x == 32) clk[blockIdx.x] = (unsigned long long)(t1 - t0);
}

// ---- the same gather with 16-byte cp.async copies (4 producer warps, thread = (piece, rows r0 + 16 i)), as k_gather_gemm_ss<3> ----
template <int PW, bool ZST = false, int ARR = 1>   // ARR: one of ARR lanes arrives on the slot barrier (timing experiment only when > 1)
   // producer warps: 4 (as k_gather_gemm_ss<3>) or 8 / 16; ZST: missing rows zeroed by a plain store instead of a zero-size cp.async
__global__ void __launch_bounds__(PW * 32 + 32) k_cpasync(const unsigned char* __restrict__ a, int64_t rows, int row_bytes, const int* __restrict__ idx,
                                                 int64_t n_stages, int col_pieces, unsigned long long* clk, int slots) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) unsigned long long full[kSlots], empty[kSlots];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
  if (tid == 0) {
    for (int s = 0; s < kSlots; ++s) { mbar_init(smem_u32(&full[s]), PW * 32 / ARR); mbar_init(smem_u32(&empty[s]), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  int* s_idx = reinterpret_cast<int*>(smem + (size_t)kSlots * kRows * kRowBytes);
  const int my_stages = (int)((n_stages - blockIdx.x + gridDim.x - 1) / gridDim.x);
  for (int i = tid; i < my_stages * kRows; i += blockDim.x) s_idx[i] = idx[((int64_t)blockIdx.x + (int64_t)(i / kRows) * gridDim.x) * kRows + i % kRows];
  __syncthreads();
  const long long t0 = clock64();
  int s = 0; uint32_t ph = 0;
  uint32_t acc = 0;
  int ls = 0;
  for (int64_t st = blockIdx.x; st < n_stages; st += gridDim.x, ++ls) {
    const uint32_t slot = smem_u32(smem) + (uint32_t)s * kRows * kRowBytes;
    if (warp < PW) {
      if (lane == 0) mbar_wait(smem_u32(&empty[s]), ph ^ 1u);
      __syncwarp();
      const int piece = tid & 7, r0 = tid >> 3;
      const int col = (int)(st % col_pieces) * kRowBytes + piece * 16;
#pragma unroll
      for (int i = 0; i < 32 / PW; ++i) {
        const int row = r0 + 4 * PW * i;
        const int src = s_idx[ls * kRows + row];
        const bool ok = (unsigned)src < (unsigned)rows;
        const unsigned char* g = a + (size_t)(ok ? src : 0) * row_bytes + col;
        const uint32_t dst = slot + (uint32_t)row * kRowBytes + (uint32_t)((piece ^ (row & 7)) << 4);
        if (ZST) {
          if (ok) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(g) : "memory");
          else asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(dst), "r"(0) : "memory");
        } else {
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(g), "r"(ok ? 16 : 0) : "memory");
        }
      }
      if (ZST) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      if (ARR == 1 || (tid % ARR) == 0) asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&full[s])) : "memory");
    } else {
      if (lane == 0) mbar_wait(smem_u32(&full[s]), ph);
      __syncwarp();
      for (int rr = lane; rr < kRows; rr += 32) acc += *reinterpret_cast<const uint32_t*>(smem + (size_t)s * kRows * kRowBytes + rr * kRowBytes);
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&empty[s]));
    }
    if (++s == slots) { s = 0; ph ^= 1u; }
  }
  const long long t1 = clock64();
  if (acc == 0x12345678u && clk == nullptr) printf("x");
  if (tid == PW * 32) clk[blockIdx.x] = (unsigned long long)(t1 - t0);
}

int main(int argc, char** argv) {
  const int zero_pct = argc > 1 ? atoi(argv[1]) : 40;   // share of "no neighbour" rows
  const int64_t M = 160000;            // rows of the operand matrix (the frame's level-0 voxels)
  const int row_bytes = 512;           // 128 channels, pre-split format
  const int col_pieces = row_bytes / kRowBytes;
  const int64_t n_stages = 148 * kStagesPerCta;
  std::vector<unsigned char> h((size_t)M * row_bytes);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (unsigned char)((i * 2654435761u) >> 13);
  std::vector<int> hidx((size_t)n_stages * kRows);
  uint32_t seed = 12345;
  for (auto& v : hidx) { seed = seed * 1664525u + 1013904223u; v = (int)((seed >> 8) % 100) < zero_pct ? -1 : (int)((seed >> 4) % M); }
  unsigned char* d_a; int* d_idx; unsigned long long* d_clk; uint32_t* d_check;
  CK(cudaMalloc(&d_a, h.size())); CK(cudaMemcpy(d_a, h.data(), h.size(), cudaMemcpyHostToDevice));
  CK(cudaMalloc(&d_idx, hidx.size() * 4)); CK(cudaMemcpy(d_idx, hidx.data(), hidx.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&d_clk, 148 * 8)); CK(cudaMalloc(&d_check, (size_t)148 * kRows * kRowBytes));
  // tensor map: [M rows][row_bytes] u8, box = 128 bytes x 1 row (gather4 takes four such rows), SWIZZLE_128B, zero fill out of range
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                               const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  EncodeFn encode = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres));
  if (!encode || qres != cudaDriverEntryPointSuccess) { printf("cuTensorMapEncodeTiled not available\n"); return 1; }
  alignas(64) CUtensorMap map;
  const cuuint64_t gdim[2] = {(cuuint64_t)row_bytes, (cuuint64_t)M};
  const cuuint64_t gstride[1] = {(cuuint64_t)row_bytes};
  const cuuint32_t estride[2] = {1, 1};
  int box_rows_ok = -1;
  for (int box_rows : {1, 4}) {   // which box height does tile::gather4 want?  (undocumented in this image: try both)
    const cuuint32_t box[2] = {(cuuint32_t)kRowBytes, (cuuint32_t)box_rows};
    CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d_a, gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode (box rows %d) failed: %d\n", box_rows, (int)r); continue; }
    CK(cudaMemset(d_check, 0xEE, (size_t)148 * kRows * kRowBytes));
    CK(cudaFuncSetAttribute(k_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, kSlots * kRows * kRowBytes + kStagesPerCta * kRows * 4));
    k_tma<<<148, 64, kSlots * kRows * kRowBytes + kStagesPerCta * kRows * 4>>>(map, d_idx, 148, col_pieces, d_clk, d_check, 4);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("gather4 with box rows %d: %s\n", box_rows, cudaGetErrorString(e)); return 1; }
    std::vector<uint32_t> got((size_t)148 * kRows * kRowBytes / 4);
    CK(cudaMemcpy(got.data(), d_check, got.size() * 4, cudaMemcpyDeviceToHost));
    // host model: slot row i = gathered row idx[i], 16-byte piece j at position j ^ (i & 7); out-of-range row = zeros
    size_t bad = 0;
    for (int cta = 0; cta < 148; ++cta)
      for (int i = 0; i < kRows; ++i) {
        const int src = hidx[(size_t)cta * kRows + i];
        const int col = (cta % col_pieces) * kRowBytes;
        for (int j = 0; j < 8; ++j)
          for (int b = 0; b < 16; ++b) {
            const unsigned char want = src < 0 ? 0 : h[(size_t)src * row_bytes + col + j * 16 + b];
            const unsigned char have = reinterpret_cast<const unsigned char*>(got.data())[((size_t)cta * kRows + i) * kRowBytes + ((j ^ (i & 7)) << 4) + b];
            bad += want != have;
          }
      }
    printf("gather4, box rows %d: %zu of %zu bytes differ from the host model (rows in issue order, SWIZZLE_128B, zero fill)\n", box_rows, bad,
           (size_t)148 * kRows * kRowBytes);
    if (bad == 0) { box_rows_ok = box_rows; break; }
  }
  if (box_rows_ok < 0) { printf("no box height reproduced the model: layout of tile::gather4 not understood\n"); return 0; }
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  std::vector<unsigned long long> clk(148);
  for (int slots : {4, 6}) {
    for (int rep = 0; rep < 2; ++rep) {
      CK(cudaEventRecord(e0));
      k_tma<<<148, 64, kSlots * kRows * kRowBytes + kStagesPerCta * kRows * 4>>>(map, d_idx, n_stages, col_pieces, d_clk, nullptr, slots);
      CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
    }
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    CK(cudaMemcpy(clk.data(), d_clk, 148 * 8, cudaMemcpyDeviceToHost));
    double mx = 0; for (auto c : clk) mx = c > mx ? c : mx;
    const double bytes = (double)n_stages * kRows * kRowBytes * (1.0 - zero_pct / 100.0);
    printf("TMA gather4 (1 issuing warp, %d slots): %.3f ms, %.1f B/clk/SM of real rows (%.0f clk per 128-row stage)\n", slots, ms, bytes / 148 / mx,
           mx / (n_stages / 148.0));
    auto run_cp = [&](auto kern, int pw) {
      CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSlots * kRows * kRowBytes + kStagesPerCta * kRows * 4));
      for (int rep = 0; rep < 2; ++rep) {
        CK(cudaEventRecord(e0));
        kern<<<148, pw * 32 + 32, kSlots * kRows * kRowBytes + kStagesPerCta * kRows * 4>>>(d_a, M, row_bytes, d_idx, n_stages, col_pieces, d_clk, slots);
        CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
      }
      CK(cudaEventElapsedTime(&ms, e0, e1));
      CK(cudaMemcpy(clk.data(), d_clk, 148 * 8, cudaMemcpyDeviceToHost));
      mx = 0; for (auto c : clk) mx = c > mx ? c : mx;
      printf("cp.async 16 B  (%d issuing warps, %d slots, %d %% zero fill): %.3f ms, %.0f clk per 128-row stage\n", pw, slots, zero_pct, ms,
             mx / (n_stages / 148.0));
    };
    run_cp(k_cpasync<4>, 4);
    run_cp(k_cpasync<8>, 8);
    run_cp(k_cpasync<16>, 16);
    printf("  missing rows by plain zero stores:\n");
    run_cp(k_cpasync<4, true>, 4);
    run_cp(k_cpasync<8, true>, 8);
    printf("  one arrival per 8 / 32 lanes (completion of the other lanes' copies NOT tracked: timing only):\n");
    run_cp(k_cpasync<4, false, 8>, 4);
    run_cp(k_cpasync<4, false, 32>, 4);
    run_cp(k_cpasync<8, false, 32>, 8);
  }
  return 0;
}
