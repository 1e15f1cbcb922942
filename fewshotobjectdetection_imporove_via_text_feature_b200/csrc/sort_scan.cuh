// Block-level sort / scan helpers shared by detect_post.cu (D3) and rpn_select.cu.
#pragma once
#include "common.cuh"

namespace b200 {

__device__ __forceinline__ int block_exclusive_scan_1024(int v, int* s_warp /*[33]*/, int* total) {
  // blockDim.x == 1024.  returns exclusive prefix of v over threads; *total = block sum
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int w = s_warp[lane];
    int winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    s_warp[lane] = winc - w;
    if (lane == 31) s_warp[32] = winc;
  }
  __syncthreads();
  const int res = s_warp[warp] + inc - v;
  *total = s_warp[32];
  __syncthreads();
  return res;
}

__device__ __forceinline__ uint32_t desc_key(float s) {
  // monotone map float -> uint32 such that larger float => SMALLER key (ascending sort == descending score);
  // -0 and +0 share a key, as they compare equal in the reference's sorts
  uint32_t u = s == 0.f ? 0u : __float_as_uint(s);
  u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  return ~u;
}

// block-wide bitonic sort (ascending) of n2 (power of two) 64-bit keys in shared or global memory
__device__ inline void bitonic_sort_u64(unsigned long long* keys, int n2) {
  for (int k = 2; k <= n2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < (n2 >> 1); t += blockDim.x) {
        const int i = ((t / j) * (j << 1)) + (t % j);
        const int l = i + j;
        const bool asc = (i & k) == 0;
        const unsigned long long a = keys[i], b = keys[l];
        if ((a > b) == asc) { keys[i] = b; keys[l] = a; }
      }
      __syncthreads();
    }
  }
}

__device__ __forceinline__ int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

}  // namespace b200
