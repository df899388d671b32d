// Shared helpers for the gnnb200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/gnnb200.h"

#define GNNB200_CHECK_CUDA(expr)                 \
  do {                                           \
    cudaError_t _e = (expr);                     \
    if (_e != cudaSuccess) return (int)_e;       \
  } while (0)

// Every launch is followed by this: reports launch-configuration errors without synchronising.  cudaGetLastError (not
// Peek): a reported error is cleared, so one bad launch does not make every later call of the process fail as well.
#define GNNB200_LAUNCH_CHECK()                   \
  do {                                           \
    cudaError_t _e = cudaGetLastError();         \
    if (_e != cudaSuccess) return (int)_e;       \
  } while (0)

namespace gnnb200 {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Bump allocator over the caller's workspace; with base == nullptr it only measures.
struct Workspace {
  char* base;
  size_t used;
  explicit Workspace(void* p) : base(static_cast<char*>(p)), used(0) {}
  template <typename T>
  T* take(size_t count) {
    used = align_up(used, 256);
    T* p = base ? reinterpret_cast<T*>(base + used) : nullptr;
    used += count * sizeof(T);
    return p;
  }
  size_t bytes() const { return align_up(used, 256); }
};

// 128-bit read-only load that does not pollute L1 (streamed neighbour rows).
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}

__device__ __forceinline__ float4 f4_add(float4 a, float4 b) {
  return make_float4(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y), __fadd_rn(a.z, b.z), __fadd_rn(a.w, b.w));
}
__device__ __forceinline__ float4 f4_scale(float s, float4 b) {
  return make_float4(__fmul_rn(s, b.x), __fmul_rn(s, b.y), __fmul_rn(s, b.z), __fmul_rn(s, b.w));
}

}  // namespace gnnb200
