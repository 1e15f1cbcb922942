// Shared helpers for the b200roi kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/b200roi.h"

namespace b200 {

// thread-local last-error text, readable through b200_last_error()
void set_error(const char* fmt, ...);

#define B200_CHECK_ARG(cond, ...)                         \
  do {                                                    \
    if (!(cond)) {                                        \
      ::b200::set_error(__VA_ARGS__);                     \
      return B200_ERR_INVALID;                            \
    }                                                     \
  } while (0)

#define B200_CUDA_LAUNCH_CHECK(what)                                                   \
  do {                                                                                 \
    cudaError_t e__ = cudaGetLastError();                                              \
    if (e__ != cudaSuccess) {                                                          \
      ::b200::set_error("%s: CUDA launch failed: %s", what, cudaGetErrorString(e__));  \
      return B200_ERR_CUDA;                                                            \
    }                                                                                  \
  } while (0)

#define B200_CUDA_CALL(expr)                                                            \
  do {                                                                                  \
    cudaError_t e__ = (expr);                                                           \
    if (e__ != cudaSuccess) {                                                           \
      ::b200::set_error("%s failed: %s", #expr, cudaGetErrorString(e__));               \
      return B200_ERR_CUDA;                                                             \
    }                                                                                   \
  } while (0)

constexpr int kNumSMs = 148;

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// 16-byte streaming store / read-only vector load
__device__ __forceinline__ void st_stream_f4(float4* p, float4 v) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ float4 ldg_f4(const float4* p) { return __ldg(p); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace b200
