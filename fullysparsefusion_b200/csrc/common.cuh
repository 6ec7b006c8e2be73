// common.cuh — shared helpers for libfsf_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <algorithm>

#include "../../include/fsf_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libfsf_b200 is written for sm_100a (B200) only"
#endif

namespace fsfb {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs; grids are sized in multiples of this

void set_error(const char* fmt, ...);
void count_launch();

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Bump allocator over a caller-provided workspace (all sub-buffers 256 B aligned).
struct Workspace {
  char* base;
  size_t size;
  size_t used;
  Workspace(void* p, size_t n) : base(static_cast<char*>(p)), size(n), used(0) {}
  template <typename T>
  T* take(size_t count) {
    size_t bytes = align_up(count * sizeof(T), 256);
    if (base == nullptr || used + bytes > size) {
      used += bytes;  // keep counting so the caller can report the need
      return nullptr;
    }
    T* p = reinterpret_cast<T*>(base + used);
    used += bytes;
    return p;
  }
  bool ok() const { return base != nullptr && used <= size; }
};

}  // namespace fsfb

#define FSFB_CHECK_ARG(cond, ...)        \
  do {                                   \
    if (!(cond)) {                       \
      fsfb::set_error(__VA_ARGS__);      \
      return FSFB_ERR_BADARG;            \
    }                                    \
  } while (0)

#define FSFB_CUDA(expr)                                                          \
  do {                                                                           \
    cudaError_t e__ = (expr);                                                    \
    if (e__ != cudaSuccess) {                                                    \
      fsfb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__),  \
                      __FILE__, __LINE__);                                       \
      return FSFB_ERR_CUDA;                                                      \
    }                                                                            \
  } while (0)

// Launch + count + check.  Usage: FSFB_LAUNCH(kernel, grid, block, smem, stream, args...)
#define FSFB_LAUNCH(kernel, grid, block, smem, stream, ...)                       \
  do {                                                                            \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                   \
    fsfb::count_launch();                                                         \
    cudaError_t e__ = cudaGetLastError();                                         \
    if (e__ != cudaSuccess) {                                                     \
      fsfb::set_error("launch %s failed: %s (%s:%d)", #kernel,                    \
                      cudaGetErrorString(e__), __FILE__, __LINE__);               \
      return FSFB_ERR_CUDA;                                                       \
    }                                                                             \
  } while (0)

namespace fsfb {

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ unsigned lanemask_lt() {
  unsigned m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

// Streaming (read-once) 128-bit load that does not allocate in L1.
__device__ __forceinline__ float4 ldg_stream_f4(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}
__device__ __forceinline__ float ldg_stream_f1(const float* p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
// Streaming 128-bit store (write-once outputs).
__device__ __forceinline__ void stg_stream_f4(float4* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x),
               "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

}  // namespace fsfb
