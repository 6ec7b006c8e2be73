// gemm_simt.cu — weight pre-packing, the standalone row norm/activation kernel, and the CUDA-core
// fp32 cross-check of the gather-GEMM contract (tests only; the product path is gemm_tc.cu).
#include <cuda_fp16.h>

#include "gemm_common.cuh"

namespace fsfb {

// ---- prepack -------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    k_gemm_prepack(const float* __restrict__ w, GemmShape S, unsigned char* __restrict__ packed) {
  // one thread per packed float position (hi and lo written together)
  const int kc = S.kc();
  const int64_t per_k = (int64_t)S.n_pad() * kc * kGemmKChunk;
  const int64_t total = per_k * S.koff;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(t / per_k);
    int64_t r = t - (int64_t)k * per_k;
    const int n = (int)(r / (kc * kGemmKChunk));
    const int i = (int)(r - (int64_t)n * kc * kGemmKChunk);  // padded cin index
    const int kchunk = i / kGemmKChunk, j = i % kGemmKChunk;
    const int nt = n / kGemmNTile, nl = n % kGemmNTile;
    const float v = (n < S.cout && i < S.cin) ? w[((int64_t)k * S.cout + n) * S.cin + i] : 0.f;
    const float hi = tf32_hi(v);
    const float lo = tf32_lo(v, hi);
    unsigned char* blk = packed + S.block_offset(nt, k, kchunk);
    const uint32_t off = sw128_offset(nl, j);
    *reinterpret_cast<float*>(blk + off) = hi;
    *reinterpret_cast<float*>(blk + (size_t)S.n_w(nt) * 128 + off) = lo;
  }
}

// fp16-split copy (FSFB_GEMM_F16=1): same walk, one 128-byte row per (n, K chunk) = [hi 32 halves | lo * 2048 32 halves]
__global__ void __launch_bounds__(256)
    k_gemm_prepack_f16(const float* __restrict__ w, GemmShape S, unsigned char* __restrict__ packed) {
  const int kc = S.kc();
  const int64_t per_k = (int64_t)S.n_pad() * kc * kGemmKChunk;
  const int64_t total = per_k * S.koff;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(t / per_k);
    int64_t r = t - (int64_t)k * per_k;
    const int n = (int)(r / (kc * kGemmKChunk));
    const int i = (int)(r - (int64_t)n * kc * kGemmKChunk);
    const int kchunk = i / kGemmKChunk, j = i % kGemmKChunk;
    const int nt = n / kGemmNTile, nl = n % kGemmNTile;
    const float v = (n < S.cout && i < S.cin) ? w[((int64_t)k * S.cout + n) * S.cin + i] : 0.f;
    const __half hi = __float2half_rn(v);
    const __half lo = __float2half_rn((v - __half2float(hi)) * kF16LoScale);
    unsigned char* blk = packed + S.f16_block_offset(nt, k, kchunk);
    *reinterpret_cast<__half*>(blk + sw128_offset_f16(nl, j)) = hi;
    *reinterpret_cast<__half*>(blk + sw128_offset_f16(nl, 32 + j)) = lo;
  }
}

// ---- row norm + act -------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    k_rownorm_act(const float* __restrict__ x, int64_t rows, int c, int64_t x_stride, Epilogue E,
                  float* __restrict__ out, int64_t out_stride) {
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < rows; r += warps)
    warp_row_epilogue(x + r * x_stride, c, E, r, out + r * out_stride);
}

// ---- fp32 CUDA-core gather-GEMM (cross-check) -------------------------------------------------
// One warp per output row; lanes own output channels.  Pre-epilogue sums go to `out`, then the
// same warp applies the row epilogue in place.
__global__ void __launch_bounds__(256)
    k_gather_gemm_simt(const float* __restrict__ a, int64_t a_rows, int cin, int64_t a_stride,
                       const int32_t* __restrict__ nbr, int koff, int64_t rows,
                       const float* __restrict__ w, int cout, Epilogue E, float* __restrict__ out,
                       int64_t out_stride) {
  const int lane = lane_id();
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < rows; r += warps) {
    float* o = out + r * out_stride;
    for (int c0 = 0; c0 < cout; c0 += 128) {
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      for (int k = 0; k < koff; ++k) {
        const int64_t src = nbr ? (int64_t)nbr[(int64_t)k * rows + r] : r;
        if (src < 0 || src >= a_rows) continue;  // warp-uniform
        const float* ar = a + src * a_stride;
        const float* wk = w + (int64_t)k * cout * cin;
        for (int i = 0; i < cin; ++i) {
          const float av = __ldg(ar + i);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int c = c0 + q * 32 + lane;
            if (c < cout) acc[q] = fmaf(av, __ldg(wk + (int64_t)c * cin + i), acc[q]);
          }
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int c = c0 + q * 32 + lane;
        if (c < cout) o[c] = acc[q];
      }
    }
    __syncwarp();
    warp_row_epilogue(o, cout, E, r, o);
    __syncwarp();
  }
}

}  // namespace fsfb

extern "C" {

int fsfb_gemm_prepack_bytes(int koff, int cin, int cout, size_t* bytes) {
  using namespace fsfb;
  FSFB_CHECK_ARG(bytes && koff >= 1 && cin >= 1 && cout >= 1, "gemm_prepack_bytes: bad argument");
  GemmShape S{koff, cin, cout};
  *bytes = gemm_f16_enabled() ? S.total_bytes_f16() : S.total_bytes();
  return FSFB_OK;
}

int fsfb_gemm_prepack(const float* w, int koff, int cin, int cout, void* packed, void* stream) {
  using namespace fsfb;
  FSFB_CHECK_ARG(w && packed && koff >= 1 && cin >= 1 && cout >= 1, "gemm_prepack: bad argument");
  FSFB_CHECK_ARG(((uintptr_t)packed & 127) == 0, "gemm_prepack: packed buffer must be 128-byte aligned");
  GemmShape S{koff, cin, cout};
  const int64_t total = (int64_t)S.n_pad() * S.kc() * kGemmKChunk * koff;
  const int grid = (int)std::min<int64_t>(ceil_div(total, 256), (int64_t)kNumSMs * 16);
  FSFB_LAUNCH(k_gemm_prepack, grid, 256, 0, (cudaStream_t)stream, w, S, (unsigned char*)packed);
  if (gemm_f16_enabled()) FSFB_LAUNCH(k_gemm_prepack_f16, grid, 256, 0, (cudaStream_t)stream, w, S, (unsigned char*)packed);
  return FSFB_OK;
}

int fsfb_rownorm_act(const float* x, int64_t rows, int c, int64_t x_stride, const float* bias, int norm,
                     const float* norm_w, const float* norm_b, float eps, const float* residual,
                     int64_t residual_stride, int act, float* out, int64_t out_stride, void* stream) {
  using namespace fsfb;
  int rc = check_epilogue(c, bias, norm, norm_w, norm_b, act, "rownorm_act");
  if (rc != FSFB_OK) return rc;
  FSFB_CHECK_ARG(rows >= 0 && x_stride >= c && out_stride >= c, "rownorm_act: bad rows/stride");
  if (rows == 0) return FSFB_OK;
  FSFB_CHECK_ARG(x && out, "rownorm_act: null pointer");
  Epilogue E{bias, norm, norm_w, norm_b, eps, residual, residual_stride, act};
  const int grid = (int)std::min<int64_t>(ceil_div(rows, 8), (int64_t)kNumSMs * 16);
  FSFB_LAUNCH(k_rownorm_act, grid, 256, 0, (cudaStream_t)stream, x, rows, c, x_stride, E, out, out_stride);
  return FSFB_OK;
}

int fsfb_gather_gemm_simt(const float* a, int64_t a_rows, int cin, int64_t a_stride, const int32_t* nbr,
                          int koff, int64_t rows, const float* w, int cout, const float* bias, int norm,
                          const float* norm_w, const float* norm_b, float eps, const float* residual,
                          int64_t residual_stride, int act, float* out, int64_t out_stride, void* stream) {
  using namespace fsfb;
  int rc = check_epilogue(cout, bias, norm, norm_w, norm_b, act, "gather_gemm_simt");
  if (rc != FSFB_OK) return rc;
  FSFB_CHECK_ARG(rows >= 0 && a_rows >= 0 && cin >= 1 && koff >= 1 && a_stride >= cin && out_stride >= cout,
                 "gather_gemm_simt: bad shape");
  FSFB_CHECK_ARG(nbr || koff == 1, "gather_gemm_simt: koff > 1 needs a neighbour table");
  if (rows == 0) return FSFB_OK;
  FSFB_CHECK_ARG(w && out && (a || a_rows == 0), "gather_gemm_simt: null pointer");
  Epilogue E{bias, norm, norm_w, norm_b, eps, residual, residual_stride, act};
  const int grid = (int)std::min<int64_t>(ceil_div(rows, 8), (int64_t)kNumSMs * 16);
  FSFB_LAUNCH(k_gather_gemm_simt, grid, 256, 0, (cudaStream_t)stream, a, a_rows, cin, a_stride, nbr, koff,
              rows, w, cout, E, out, out_stride);
  return FSFB_OK;
}

}  // extern "C"
