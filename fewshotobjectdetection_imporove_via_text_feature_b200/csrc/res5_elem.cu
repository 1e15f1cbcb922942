// P2 glue: the elementwise passes around the frozen res5 convolutions (defrcn/modeling/roi_heads/roi_heads.py:313-344
// `_build_res5_block` / `_shared_roi_transform`, and the spatial mean at :1109), forward and backward.
//
// res5 itself stays on cuDNN/cuBLAS.  What autograd adds around the data-gradient convolutions is
// pure HBM traffic over (R, 4, 4, 2048) bf16 tensors (268 MB each at R = 4096), and torch spends one kernel per
// algebraic step: mean backward = scale (fp32) + cast + expand-copy + ReLU mask (4 passes, 0.81 ms in the round-1
// launch list); residual fan-in = add + ReLU mask (2 passes, 0.22 ms per block).  Each group is one pass here:
//   spatial_mean            pooled[r,c]  = 1/HW * sum_p out[r,p,c]                       (bf16 NHWC in, fp32 out)
//   mean_bwd_relu_mask      g[r,p,c]     = out[r,p,c] > 0 ? gpooled[r,c] / HW : 0         (fp32 in, bf16 NHWC out)
//   add_relu_mask           y[i]         = ref[i] > 0 ? bf16(a[i] + b[i]) : 0             (residual fan-in + ReLU bwd)
// All three are bound by HBM bandwidth: 16-byte accesses, consecutive lanes on consecutive channels.
//
// 1-bit ReLU masks.  Every one of these backward passes reads a whole activation tensor only to test `> 0`: a quarter of
// add_relu_mask's bytes, half of mean_bwd_relu_mask's, a third of each threshold_backward's.  The `_bits` variants take
// the mask as one byte per 8 consecutive elements (bit k <-> element 8 i + k, the same keep rule) instead: 1/16 of the
// bytes.  The masks are written where the activation is read anyway (spatial_mean) or by pack_relu_bits, which the
// forward launches on a side stream under the compute-bound convolutions.
#include "common.cuh"

namespace b200 {

__device__ __forceinline__ void unpack8(const uint4 t, float* f) {
  f[0] = __uint_as_float(t.x << 16); f[1] = __uint_as_float(t.x & 0xffff0000u);
  f[2] = __uint_as_float(t.y << 16); f[3] = __uint_as_float(t.y & 0xffff0000u);
  f[4] = __uint_as_float(t.z << 16); f[5] = __uint_as_float(t.z & 0xffff0000u);
  f[6] = __uint_as_float(t.w << 16); f[7] = __uint_as_float(t.w & 0xffff0000u);
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&h);
}
// per 16-bit half: 0xffff where the gradient passes the ReLU, i.e. NOT (x <= 0) — aten's threshold_backward rule, under
// which a NaN activation lets the gradient through: keep iff NaN, or sign clear and magnitude non-zero
__device__ __forceinline__ uint32_t positive_mask2(uint32_t v) {
  uint32_t m = 0;
  const uint32_t lo = v & 0xffffu, hi = v >> 16;
  if ((lo & 0x7fffu) > 0x7f80u || (lo != 0 && lo < 0x8000u)) m |= 0xffffu;
  if ((hi & 0x7fffu) > 0x7f80u || (hi != 0 && hi < 0x8000u)) m |= 0xffff0000u;
  return m;
}


// bit k of the result <-> element k of the 8 bf16 in `v` passes the ReLU backward (same rule as positive_mask2)
__device__ __forceinline__ uint32_t keep_bits8(const uint4 v) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
  uint32_t bits = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t m = positive_mask2(w[j]);
    bits |= ((m & 1u) | ((m >> 16 & 1u) << 1)) << (2 * j);
  }
  return bits;
}
// expand bits (2 j, 2 j + 1) of a mask byte to the 0xffff-per-half mask of word j
__device__ __forceinline__ uint32_t expand_mask2(uint32_t bits, int j) {
  const uint32_t b = bits >> (2 * j);
  return ((b & 1u) ? 0xffffu : 0u) | ((b & 2u) ? 0xffff0000u : 0u);
}

// thread = (roi, 8 channels); HW pixels walked serially (HW = 16 for res5's 4x4 output)
__global__ void __launch_bounds__(256)
spatial_mean_kernel(const uint4* __restrict__ x, float* __restrict__ pooled, int ld_pooled, int R, int HW, int C8, float inv,
                    uint8_t* __restrict__ bits) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)R * C8) return;
  const int r = (int)(i / C8), c8 = (int)(i - (size_t)r * C8);
  const uint4* src = x + (size_t)r * HW * C8 + c8;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
  for (int p = 0; p < HW; ++p) {
    float f[8];
    const uint4 v = __ldcs(src + (size_t)p * C8);
    unpack8(v, f);
    if (bits) bits[((size_t)r * HW + p) * C8 + c8] = (uint8_t)keep_bits8(v);
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] += f[k];
  }
  float4* dst = reinterpret_cast<float4*>(pooled + (size_t)r * ld_pooled + c8 * 8);
  dst[0] = make_float4(acc[0] * inv, acc[1] * inv, acc[2] * inv, acc[3] * inv);
  dst[1] = make_float4(acc[4] * inv, acc[5] * inv, acc[6] * inv, acc[7] * inv);
}

__global__ void __launch_bounds__(256)
mean_bwd_relu_mask_kernel(const float* __restrict__ gpooled, int ld_g, const uint4* __restrict__ out,
                          const uint8_t* __restrict__ bits, uint4* __restrict__ g, int R, int HW, int C8, float inv,
                          int interleaved) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)R * C8) return;
  const int r = (int)(i / C8), c8 = (int)(i - (size_t)r * C8);
  const float4* gp = reinterpret_cast<const float4*>(gpooled + (size_t)r * ld_g + c8 * 8);
  const float4 a = __ldg(gp), b = __ldg(gp + 1);
  uint4 v;                                                // bf16(gpooled / HW), the value torch's cast would produce
  v.x = pack2(a.x * inv, a.y * inv); v.y = pack2(a.z * inv, a.w * inv);
  v.z = pack2(b.x * inv, b.y * inv); v.w = pack2(b.z * inv, b.w * inv);
  uint4* dst = g + (size_t)r * HW * C8 + c8;
  if (bits && interleaved) {
    // masks written by the GEMM epilogue (csrc/gemm2_tcgen05.cu): per 32 columns one word, bit j <-> column 2j, bit 16 + j
    // <-> column 2j + 1; this thread's 8 columns are nibble k = c8 % 4 of each half
    const uint32_t* mw = reinterpret_cast<const uint32_t*>(bits) + (size_t)r * HW * (C8 / 4) + (c8 >> 2);
    const int sh = 4 * (c8 & 3);
#pragma unroll 4
    for (int p = 0; p < HW; ++p) {
      const uint32_t word = __ldg(mw + (size_t)p * (C8 / 4));
      const uint32_t lo = word >> sh, hi = word >> (16 + sh);
      uint4 w;
      w.x = v.x & (((lo & 1u) ? 0xffffu : 0u) | ((hi & 1u) ? 0xffff0000u : 0u));
      w.y = v.y & (((lo & 2u) ? 0xffffu : 0u) | ((hi & 2u) ? 0xffff0000u : 0u));
      w.z = v.z & (((lo & 4u) ? 0xffffu : 0u) | ((hi & 4u) ? 0xffff0000u : 0u));
      w.w = v.w & (((lo & 8u) ? 0xffffu : 0u) | ((hi & 8u) ? 0xffff0000u : 0u));
      dst[(size_t)p * C8] = w;
    }
    return;
  }
  if (bits) {
    const uint8_t* mb = bits + (size_t)r * HW * C8 + c8;
#pragma unroll 4
    for (int p = 0; p < HW; ++p) {
      const uint32_t b = mb[(size_t)p * C8];
      uint4 w;
      w.x = v.x & expand_mask2(b, 0); w.y = v.y & expand_mask2(b, 1);
      w.z = v.z & expand_mask2(b, 2); w.w = v.w & expand_mask2(b, 3);
      dst[(size_t)p * C8] = w;
    }
    return;
  }
  const uint4* src = out + (size_t)r * HW * C8 + c8;
#pragma unroll 4
  for (int p = 0; p < HW; ++p) {
    const uint4 o = __ldcs(src + (size_t)p * C8);
    uint4 w;
    w.x = v.x & positive_mask2(o.x); w.y = v.y & positive_mask2(o.y);
    w.z = v.z & positive_mask2(o.z); w.w = v.w & positive_mask2(o.w);
    dst[(size_t)p * C8] = w;
  }
}

__global__ void __launch_bounds__(256)
add_relu_mask_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b, const uint4* __restrict__ ref,
                     const uint8_t* __restrict__ bits, uint4* __restrict__ y, size_t n8) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
    uint4 v = __ldcs(a + i);
    if (b) {
      float fa[8], fb[8];
      unpack8(v, fa);
      unpack8(__ldcs(b + i), fb);
      v.x = pack2(fa[0] + fb[0], fa[1] + fb[1]); v.y = pack2(fa[2] + fb[2], fa[3] + fb[3]);
      v.z = pack2(fa[4] + fb[4], fa[5] + fb[5]); v.w = pack2(fa[6] + fb[6], fa[7] + fb[7]);
    }
    if (bits) {
      const uint32_t m = bits[i];
      v.x &= expand_mask2(m, 0); v.y &= expand_mask2(m, 1); v.z &= expand_mask2(m, 2); v.w &= expand_mask2(m, 3);
    } else if (ref) {
      const uint4 o = __ldcs(ref + i);
      v.x &= positive_mask2(o.x); v.y &= positive_mask2(o.y); v.z &= positive_mask2(o.z); v.w &= positive_mask2(o.w);
    }
    y[i] = v;
  }
}

__global__ void __launch_bounds__(256)
pack_relu_bits_kernel(const uint4* __restrict__ x, uint8_t* __restrict__ bits, size_t n8) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) bits[i] = (uint8_t)keep_bits8(__ldg(x + i));
}

}  // namespace b200

using namespace b200;

extern "C" int b200_spatial_mean(const void* x_bf16, float* pooled, int ld_pooled, int R, int HW, int C, b200_stream_t stream) {
  B200_CHECK_ARG(x_bf16 && pooled, "spatial_mean: null tensor");
  B200_CHECK_ARG(R >= 0 && HW > 0 && C > 0 && C % 8 == 0 && ld_pooled % 4 == 0 && ld_pooled >= C,
                 "spatial_mean: need C %% 8 == 0 and ld_pooled %% 4 == 0");
  B200_CHECK_ARG((((uintptr_t)x_bf16 | (uintptr_t)pooled) & 15) == 0, "spatial_mean: pointers must be 16-byte aligned");
  if (R == 0) return B200_OK;
  const size_t n = (size_t)R * (C / 8);
  spatial_mean_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const uint4*)x_bf16, pooled, ld_pooled, R,
                                                                                   HW, C / 8, 1.0f / (float)HW, nullptr);
  B200_CUDA_LAUNCH_CHECK("spatial_mean");
  return B200_OK;
}

extern "C" int b200_spatial_mean_bits(const void* x_bf16, float* pooled, int ld_pooled, void* relu_bits, int R, int HW, int C,
                                      b200_stream_t stream) {
  B200_CHECK_ARG(x_bf16 && pooled && relu_bits, "spatial_mean_bits: null tensor");
  B200_CHECK_ARG(R >= 0 && HW > 0 && C > 0 && C % 8 == 0 && ld_pooled % 4 == 0 && ld_pooled >= C,
                 "spatial_mean_bits: need C %% 8 == 0 and ld_pooled %% 4 == 0");
  B200_CHECK_ARG((((uintptr_t)x_bf16 | (uintptr_t)pooled) & 15) == 0, "spatial_mean_bits: pointers must be 16-byte aligned");
  if (R == 0) return B200_OK;
  const size_t n = (size_t)R * (C / 8);
  spatial_mean_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const uint4*)x_bf16, pooled, ld_pooled, R,
                                                                                   HW, C / 8, 1.0f / (float)HW, (uint8_t*)relu_bits);
  B200_CUDA_LAUNCH_CHECK("spatial_mean_bits");
  return B200_OK;
}

extern "C" int b200_pack_relu_bits(const void* x_bf16, void* relu_bits, size_t n, b200_stream_t stream) {
  B200_CHECK_ARG(x_bf16 && relu_bits, "pack_relu_bits: null tensor");
  B200_CHECK_ARG(n % 8 == 0 && ((uintptr_t)x_bf16 & 15) == 0, "pack_relu_bits: element count %% 8, 16-byte aligned input");
  if (n == 0) return B200_OK;
  const size_t n8 = n / 8;
  const unsigned grid = (unsigned)min((n8 + 255) / 256, (size_t)kNumSMs * 16);
  pack_relu_bits_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const uint4*)x_bf16, (uint8_t*)relu_bits, n8);
  B200_CUDA_LAUNCH_CHECK("pack_relu_bits");
  return B200_OK;
}

extern "C" int b200_mean_bwd_relu_bits(const float* gpooled, int ld_g, const void* relu_bits, void* g_bf16, int R, int HW, int C, int bit_layout,
                                       b200_stream_t stream) {
  B200_CHECK_ARG(gpooled && relu_bits && g_bf16, "mean_bwd_relu_bits: null tensor");
  B200_CHECK_ARG(R >= 0 && HW > 0 && C > 0 && C % 8 == 0 && ld_g % 4 == 0 && ld_g >= C,
                 "mean_bwd_relu_bits: need C %% 8 == 0 and ld_g %% 4 == 0");
  B200_CHECK_ARG(bit_layout == 0 || (bit_layout == 1 && C % 32 == 0 && ((uintptr_t)relu_bits & 3) == 0),
                 "mean_bwd_relu_bits: bit_layout 1 (GEMM epilogue words) needs C %% 32 == 0");
  B200_CHECK_ARG((((uintptr_t)gpooled | (uintptr_t)g_bf16) & 15) == 0, "mean_bwd_relu_bits: pointers must be 16-byte aligned");
  if (R == 0) return B200_OK;
  const size_t n = (size_t)R * (C / 8);
  mean_bwd_relu_mask_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      gpooled, ld_g, nullptr, (const uint8_t*)relu_bits, (uint4*)g_bf16, R, HW, C / 8, 1.0f / (float)HW, bit_layout);
  B200_CUDA_LAUNCH_CHECK("mean_bwd_relu_bits");
  return B200_OK;
}

extern "C" int b200_add_relu_bits(const void* a_bf16, const void* b_bf16, const void* relu_bits, void* y_bf16, size_t n,
                                  b200_stream_t stream) {
  B200_CHECK_ARG(a_bf16 && y_bf16 && relu_bits, "add_relu_bits: null tensor");
  B200_CHECK_ARG(n % 8 == 0, "add_relu_bits: element count must be a multiple of 8");
  B200_CHECK_ARG((((uintptr_t)a_bf16 | (uintptr_t)b_bf16 | (uintptr_t)y_bf16) & 15) == 0,
                 "add_relu_bits: pointers must be 16-byte aligned");
  if (n == 0) return B200_OK;
  const size_t n8 = n / 8;
  const unsigned grid = (unsigned)min((n8 + 255) / 256, (size_t)kNumSMs * 32);
  add_relu_mask_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const uint4*)a_bf16, (const uint4*)b_bf16, nullptr,
                                                               (const uint8_t*)relu_bits, (uint4*)y_bf16, n8);
  B200_CUDA_LAUNCH_CHECK("add_relu_bits");
  return B200_OK;
}

extern "C" int b200_mean_bwd_relu_mask(const float* gpooled, int ld_g, const void* out_bf16, void* g_bf16, int R, int HW, int C,
                                       b200_stream_t stream) {
  B200_CHECK_ARG(gpooled && out_bf16 && g_bf16, "mean_bwd_relu_mask: null tensor");
  B200_CHECK_ARG(R >= 0 && HW > 0 && C > 0 && C % 8 == 0 && ld_g % 4 == 0 && ld_g >= C,
                 "mean_bwd_relu_mask: need C %% 8 == 0 and ld_g %% 4 == 0");
  B200_CHECK_ARG((((uintptr_t)gpooled | (uintptr_t)out_bf16 | (uintptr_t)g_bf16) & 15) == 0,
                 "mean_bwd_relu_mask: pointers must be 16-byte aligned");
  if (R == 0) return B200_OK;
  const size_t n = (size_t)R * (C / 8);
  mean_bwd_relu_mask_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      gpooled, ld_g, (const uint4*)out_bf16, nullptr, (uint4*)g_bf16, R, HW, C / 8, 1.0f / (float)HW, 0);
  B200_CUDA_LAUNCH_CHECK("mean_bwd_relu_mask");
  return B200_OK;
}

extern "C" int b200_add_relu_mask(const void* a_bf16, const void* b_bf16, const void* ref_bf16, void* y_bf16, size_t n,
                                  b200_stream_t stream) {
  B200_CHECK_ARG(a_bf16 && y_bf16, "add_relu_mask: null tensor");
  B200_CHECK_ARG(n % 8 == 0, "add_relu_mask: element count must be a multiple of 8");
  B200_CHECK_ARG((((uintptr_t)a_bf16 | (uintptr_t)b_bf16 | (uintptr_t)ref_bf16 | (uintptr_t)y_bf16) & 15) == 0,
                 "add_relu_mask: pointers must be 16-byte aligned");
  if (n == 0) return B200_OK;
  const size_t n8 = n / 8;
  const unsigned grid = (unsigned)min((n8 + 255) / 256, (size_t)kNumSMs * 32);
  add_relu_mask_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const uint4*)a_bf16, (const uint4*)b_bf16,
                                                               (const uint4*)ref_bf16, nullptr, (uint4*)y_bf16, n8);
  B200_CUDA_LAUNCH_CHECK("add_relu_mask");
  return B200_OK;
}
