// G1 + G2: Gradient Decoupled Layer + per-channel affine, fused with the layout / dtype change the
// ROIAlign gather wants (NCHW fp32 backbone output -> NHWC fp32|bf16), forward and backward.
// Reference semantics: defrcn/modeling/meta_arch/gdl.py:6-38, call site rcnn.py:94-97.
// HBM-bound elementwise work: one read + one write of the res4 map per direction.
#include <type_traits>

#include "common.cuh"

namespace b200 {

template <typename T> __device__ __forceinline__ float ldf(const T* p, size_t i);
template <> __device__ __forceinline__ float ldf<float>(const float* p, size_t i) { return __ldg(p + i); }
template <> __device__ __forceinline__ float ldf<__nv_bfloat16>(const __nv_bfloat16* p, size_t i) {
  return __bfloat162float(p[i]);
}
template <typename T> __device__ __forceinline__ void stf(T* p, size_t i, float v);
template <> __device__ __forceinline__ void stf<float>(float* p, size_t i, float v) { p[i] = v; }
template <> __device__ __forceinline__ void stf<__nv_bfloat16>(__nv_bfloat16* p, size_t i, float v) {
  p[i] = __float2bfloat16_rn(v);
}

// Rounding follows the reference's op order exactly: (x*w) then (+b) in forward; (g*w) then (*lambda) in
// backward — separate roundings, no FMA contraction — so fp32 results are bit-identical to torch.
__device__ __forceinline__ float affine_op(float x, float wc, float mult, bool has_b, float bc) {
  const float v = __fmul_rn(__fmul_rn(x, wc), mult);
  return has_b ? __fadd_rn(v, bc) : v;
}

// same-layout elementwise: y = x * (w[c] * mult) + b[c].  inner = HW for NCHW (c = (i/inner)%C),
// inner = 1 for NHWC (c = i % C).
template <typename Tin, typename Tout>
__global__ void affine_same_layout_kernel(const Tin* __restrict__ x, const float* __restrict__ w,
                                          const float* __restrict__ b, float mult, Tout* __restrict__ y,
                                          size_t total, int C, int inner) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int c = (int)((i / (size_t)inner) % (size_t)C);
    stf<Tout>(y, i, affine_op(ldf<Tin>(x, i), w ? __ldg(w + c) : 1.f, mult, b != nullptr, b ? __ldg(b + c) : 0.f));
  }
}

// fp32 vectorised variant (16-byte accesses); requires the 4 elements to share the NCHW channel
// (inner % 4 == 0) or to be 4 consecutive NHWC channels (C % 4 == 0).
__global__ void affine_same_layout_vec4_kernel(const float4* __restrict__ x, const float* __restrict__ w,
                                               const float* __restrict__ b, float mult,
                                               float4* __restrict__ y, size_t total4, int C, int inner,
                                               int nhwc) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += stride) {
    float4 v = __ldg(x + i);
    const size_t e = i * 4;
    if (nhwc) {
      const int c = (int)(e % (size_t)C);
      float4 s = w ? *reinterpret_cast<const float4*>(w + c) : make_float4(1.f, 1.f, 1.f, 1.f);
      float4 bb = b ? *reinterpret_cast<const float4*>(b + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      v.x = affine_op(v.x, s.x, mult, b != nullptr, bb.x); v.y = affine_op(v.y, s.y, mult, b != nullptr, bb.y);
      v.z = affine_op(v.z, s.z, mult, b != nullptr, bb.z); v.w = affine_op(v.w, s.w, mult, b != nullptr, bb.w);
    } else {
      const int c = (int)((e / (size_t)inner) % (size_t)C);
      const float s = w ? __ldg(w + c) : 1.f;
      const float bb = b ? __ldg(b + c) : 0.f;
      v.x = affine_op(v.x, s, mult, b != nullptr, bb); v.y = affine_op(v.y, s, mult, b != nullptr, bb);
      v.z = affine_op(v.z, s, mult, b != nullptr, bb); v.w = affine_op(v.w, s, mult, b != nullptr, bb);
    }
    y[i] = v;
  }
}

// layout-changing variant: x viewed as [N][A][B] -> y [N][B][A] through a padded smem tile.
// a_is_channel: NCHW->NHWC (A = C, B = HW); otherwise NHWC->NCHW (A = HW, B = C).
template <typename Tin, typename Tout>
__global__ void affine_transpose_kernel(const Tin* __restrict__ x, const float* __restrict__ w,
                                        const float* __restrict__ b, float mult, Tout* __restrict__ y,
                                        int A, int B, int a_is_channel) {
  __shared__ float tile[32][33];
  const int tiles_b = (B + 31) / 32;
  const int n = blockIdx.x / tiles_b;               // batch folded into grid.x (grid.z is limited to 65535)
  const int a0 = blockIdx.y * 32, b0 = (blockIdx.x % tiles_b) * 32;
  const size_t base = (size_t)n * A * B;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int a = a0 + r, bb = b0 + threadIdx.x;
    if (a < A && bb < B) {
      const int c = a_is_channel ? a : bb;
      tile[r][threadIdx.x] = affine_op(ldf<Tin>(x, base + (size_t)a * B + bb), w ? __ldg(w + c) : 1.f, mult,
                                       b != nullptr, b ? __ldg(b + c) : 0.f);
    }
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int bb = b0 + r, a = a0 + threadIdx.x;
    if (a < A && bb < B) stf<Tout>(y, base + (size_t)bb * A + a, tile[threadIdx.x][r]);
  }
}

// Vectorised layout-changing variant: 64 x 64 tiles, 16-byte (fp32) / 8-byte (bf16) global accesses on both sides,
// 16 elements per thread.  x viewed as [N][A][B] -> y [N][B][A]; needs A % 4 == 0, B % 4 == 0 and aligned bases.
// WITH_GRAD (backward, NHWC gradient -> NCHW grad_x, so B = C): the same pass also reads the forward input x (NCHW, the
// output's addressing) and leaves per-(image, tile) partial sums of dw[c] = sum g x and db[c] = sum g for an ordered
// final reduction — the separate parameter-gradient pass re-read both maps with channel-strided accesses.
constexpr int kAT = 64;

template <typename T> struct Vec4;
template <> struct Vec4<float> {
  static __device__ __forceinline__ void load(const float* p, float* v) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  static __device__ __forceinline__ void store(float* p, const float* v) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <> struct Vec4<__nv_bfloat16> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float* v) {
    const uint2 t = __ldg(reinterpret_cast<const uint2*>(p));
    v[0] = __uint_as_float(t.x << 16); v[1] = __uint_as_float(t.x & 0xffff0000u);
    v[2] = __uint_as_float(t.y << 16); v[3] = __uint_as_float(t.y & 0xffff0000u);
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float* v) {
    const __nv_bfloat162 h0 = __floats2bfloat162_rn(v[0], v[1]), h1 = __floats2bfloat162_rn(v[2], v[3]);
    uint2 t;
    t.x = *reinterpret_cast<const uint32_t*>(&h0);
    t.y = *reinterpret_cast<const uint32_t*>(&h1);
    *reinterpret_cast<uint2*>(p) = t;
  }
};

template <typename Tin, typename Tout, bool WITH_GRAD>
__global__ void __launch_bounds__(256)
affine_tile_kernel(const Tin* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b, float mult,
                   Tout* __restrict__ y, int A, int B, int a_is_channel, const float* __restrict__ fwd_x,
                   float* __restrict__ partial) {
  __shared__ float tile[kAT][kAT + 1];
  const int tiles_b = (B + kAT - 1) / kAT;
  const int n = blockIdx.x / tiles_b;               // batch folded into grid.x (grid.z is limited to 65535)
  const int a0 = blockIdx.y * kAT, b0 = (blockIdx.x % tiles_b) * kAT;
  const size_t base = (size_t)n * A * B;
  const int t = threadIdx.x;
  // load: thread <-> (row a = t / 16 + 16 i, 4 consecutive b); raw values (the affine is applied on the way out)
  {
    const int bq = (t & 15) * 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int ar = (t >> 4) + 16 * i;
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (a0 + ar < A && b0 + bq < B) Vec4<Tin>::load(x + base + (size_t)(a0 + ar) * B + b0 + bq, v);
#pragma unroll
      for (int k = 0; k < 4; ++k) tile[ar][bq + k] = v[k];
    }
  }
  __syncthreads();
  // store: thread <-> (column b = t / 16 + 16 i, 4 consecutive a)
  const int aq = (t & 15) * 4;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int br = (t >> 4) + 16 * i;
    const int bb = b0 + br, a = a0 + aq;
    const bool ok = bb < B && a < A;
    float g[4], v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) g[k] = tile[aq + k][br];
    if (ok) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int c = a_is_channel ? a + k : bb;
        v[k] = affine_op(g[k], w ? __ldg(w + c) : 1.f, mult, b != nullptr, b ? __ldg(b + c) : 0.f);
      }
      Vec4<Tout>::store(y + base + (size_t)bb * A + a, v);
    }
    if (WITH_GRAD) {
      float gw = 0.f, gb = 0.f;
      if (ok) {
        float xv[4];
        Vec4<float>::load(fwd_x + base + (size_t)bb * A + a, xv);
#pragma unroll
        for (int k = 0; k < 4; ++k) { gw += g[k] * xv[k]; gb += g[k]; }
      }
      // the 16 lanes of a half-warp hold the tile's 64 positions of channel bb
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) {
        gw += __shfl_xor_sync(0xffffffffu, gw, o);
        gb += __shfl_xor_sync(0xffffffffu, gb, o);
      }
      if ((t & 15) == 0 && bb < B) {
        const size_t chunk = (size_t)n * gridDim.y + blockIdx.y;
        partial[(chunk * B + bb) * 2 + 0] = gw;
        partial[(chunk * B + bb) * 2 + 1] = gb;
      }
    }
  }
}

static bool affine_tile_ok(const void* x, const void* y, int A, int B) {
  return A % 4 == 0 && B % 4 == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 15) == 0;
}

template <typename Tin, typename Tout>
static int launch_affine(const void* x, const float* w, const float* b, float mult, void* y, int N, int C,
                         int H, int W, int in_layout, int out_layout, cudaStream_t st) {
  const int HW = H * W;
  const size_t total = (size_t)N * C * HW;
  if (total == 0) return B200_OK;
  if (in_layout == out_layout) {
    const int inner = in_layout == B200_NCHW ? HW : 1;
    const bool vec_ok = std::is_same<Tin, float>::value && std::is_same<Tout, float>::value &&
                        ((in_layout == B200_NCHW && HW % 4 == 0) || (in_layout == B200_NHWC && C % 4 == 0)) &&
                        ((uintptr_t)x % 16 == 0) && ((uintptr_t)y % 16 == 0) &&
                        (in_layout == B200_NCHW || ((!w || (uintptr_t)w % 16 == 0) && (!b || (uintptr_t)b % 16 == 0)));
    if (vec_ok) {
      const size_t t4 = total / 4;
      const int blocks = (int)min((size_t)kNumSMs * 8, (t4 + 255) / 256);
      affine_same_layout_vec4_kernel<<<blocks, 256, 0, st>>>((const float4*)x, w, b, mult, (float4*)y, t4, C,
                                                              inner, in_layout == B200_NHWC);
    } else {
      const int blocks = (int)min((size_t)kNumSMs * 8, (total + 255) / 256);
      affine_same_layout_kernel<Tin, Tout><<<blocks, 256, 0, st>>>((const Tin*)x, w, b, mult, (Tout*)y, total, C, inner);
    }
  } else {
    const int a_is_channel = in_layout == B200_NCHW;
    const int A = a_is_channel ? C : HW, B = a_is_channel ? HW : C;
    if (affine_tile_ok(x, y, A, B)) {
      dim3 grid(ceil_div(B, kAT) * N, ceil_div(A, kAT));
      affine_tile_kernel<Tin, Tout, false><<<grid, 256, 0, st>>>((const Tin*)x, w, b, mult, (Tout*)y, A, B, a_is_channel,
                                                                 nullptr, nullptr);
    } else {
      dim3 grid(ceil_div(B, 32) * N, ceil_div(A, 32)), block(32, 8);
      affine_transpose_kernel<Tin, Tout><<<grid, block, 0, st>>>((const Tin*)x, w, b, mult, (Tout*)y, A, B, a_is_channel);
    }
  }
  B200_CUDA_LAUNCH_CHECK("gdl_affine");
  return B200_OK;
}

int dispatch_affine(const void* x, const float* w, const float* b, float mult, void* y, int N, int C, int H, int W,
                    int in_dtype, int in_layout, int out_dtype, int out_layout, cudaStream_t st) {
  if (in_dtype == B200_F32 && out_dtype == B200_F32)
    return launch_affine<float, float>(x, w, b, mult, y, N, C, H, W, in_layout, out_layout, st);
  if (in_dtype == B200_F32 && out_dtype == B200_BF16)
    return launch_affine<float, __nv_bfloat16>(x, w, b, mult, y, N, C, H, W, in_layout, out_layout, st);
  if (in_dtype == B200_BF16 && out_dtype == B200_F32)
    return launch_affine<__nv_bfloat16, float>(x, w, b, mult, y, N, C, H, W, in_layout, out_layout, st);
  return launch_affine<__nv_bfloat16, __nv_bfloat16>(x, w, b, mult, y, N, C, H, W, in_layout, out_layout, st);
}

// ---- backward parameter gradients: deterministic two-pass reduction --------------------------------
constexpr int kRedChunks = 64;

__device__ __forceinline__ size_t map_index(int layout, int n, int c, int p, int C, int HW) {
  return layout == B200_NCHW ? ((size_t)n * C + c) * HW + p : ((size_t)n * HW + p) * C + c;
}

// block (32 channels, 8 pixel lanes); grid (ceil(C/32), kRedChunks).  partial[chunk][c][2]
template <typename Tx, typename Tg>
__global__ void affine_param_grad_partial_kernel(const Tg* __restrict__ gy, const Tx* __restrict__ x,
                                                 float* __restrict__ partial, int N, int C, int HW,
                                                 int x_layout, int g_layout) {
  __shared__ float s_gw[8][33], s_gb[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const long long total_p = (long long)N * HW;
  const long long per = (total_p + gridDim.y - 1) / gridDim.y;
  const long long p0 = per * blockIdx.y, p1 = min(total_p, p0 + per);
  float gw = 0.f, gb = 0.f;
  if (c < C) {
    for (long long q = p0 + threadIdx.y; q < p1; q += 8) {
      const int n = (int)(q / HW), p = (int)(q % HW);
      const float g = ldf<Tg>(gy, map_index(g_layout, n, c, p, C, HW));
      const float xv = ldf<Tx>(x, map_index(x_layout, n, c, p, C, HW));
      gw += g * xv;
      gb += g;
    }
  }
  s_gw[threadIdx.y][threadIdx.x] = gw;
  s_gb[threadIdx.y][threadIdx.x] = gb;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    float a = 0.f, bsum = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) { a += s_gw[k][threadIdx.x]; bsum += s_gb[k][threadIdx.x]; }
    partial[((size_t)blockIdx.y * C + c) * 2 + 0] = a;
    partial[((size_t)blockIdx.y * C + c) * 2 + 1] = bsum;
  }
}

__global__ void affine_param_grad_final_kernel(const float* __restrict__ partial, float* __restrict__ grad_w,
                                               float* __restrict__ grad_b, int C, int chunks) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float a = 0.f, b = 0.f;
  for (int k = 0; k < chunks; ++k) {
    a += partial[((size_t)k * C + c) * 2 + 0];
    b += partial[((size_t)k * C + c) * 2 + 1];
  }
  if (grad_w) grad_w[c] = a;
  if (grad_b) grad_b[c] = b;
}

}  // namespace b200

using namespace b200;

extern "C" int b200_gdl_affine_fwd(const void* x, const float* weight, const float* bias, void* y, int N, int C,
                                   int H, int W, int in_dtype, int in_layout, int out_dtype, int out_layout,
                                   b200_stream_t stream) {
  B200_CHECK_ARG(x && y, "gdl_affine_fwd: null tensor");
  B200_CHECK_ARG(N >= 0 && C > 0 && H > 0 && W > 0, "gdl_affine_fwd: bad shape");
  B200_CHECK_ARG((in_dtype | 1) == 1 && (out_dtype | 1) == 1 && (in_layout | 1) == 1 && (out_layout | 1) == 1,
                 "gdl_affine_fwd: bad dtype/layout");
  return dispatch_affine(x, weight, bias, 1.0f, y, N, C, H, W, in_dtype, in_layout, out_dtype, out_layout,
                         (cudaStream_t)stream);
}

extern "C" size_t b200_gdl_affine_bwd_workspace_bytes(int N, int C, int H, int W) {
  const size_t chunks = max((size_t)kRedChunks, (size_t)max(N, 1) * (size_t)ceil_div(H * W, kAT));
  return chunks * C * 2 * sizeof(float);
}

extern "C" int b200_gdl_affine_bwd(const void* grad_y, const void* x, const float* weight, float lambda,
                                   void* grad_x, float* grad_w, float* grad_b, int N, int C, int H, int W,
                                   int in_dtype, int in_layout, int out_dtype, int out_layout, void* workspace,
                                   size_t workspace_bytes, b200_stream_t stream) {
  B200_CHECK_ARG(grad_y, "gdl_affine_bwd: null grad_y");
  cudaStream_t st = (cudaStream_t)stream;
  const int HW_ = H * W;
  // The fine-tune path's shape: bf16 channels-last gradient in, fp32 NCHW grad_x out, parameter gradients wanted:
  // one pass does the layout change, the scaling and the per-tile partial sums of dw / db.
  if (grad_x && (grad_w || grad_b) && x && in_dtype == B200_F32 && in_layout == B200_NCHW && out_dtype == B200_BF16 &&
      out_layout == B200_NHWC && affine_tile_ok(grad_y, grad_x, HW_, C) && ((uintptr_t)x & 15) == 0 && N > 0) {
    const int chunks = N * ceil_div(HW_, kAT);
    if (!workspace || workspace_bytes < (size_t)chunks * C * 2 * sizeof(float)) {
      set_error("gdl_affine_bwd: workspace too small");
      return B200_ERR_WORKSPACE;
    }
    float* partial = (float*)workspace;
    dim3 grid(ceil_div(C, kAT) * N, ceil_div(HW_, kAT));
    affine_tile_kernel<__nv_bfloat16, float, true><<<grid, 256, 0, st>>>((const __nv_bfloat16*)grad_y, weight, nullptr, lambda,
                                                                         (float*)grad_x, HW_, C, 0, (const float*)x, partial);
    B200_CUDA_LAUNCH_CHECK("gdl_affine_bwd tile");
    affine_param_grad_final_kernel<<<ceil_div(C, 128), 128, 0, st>>>(partial, grad_w, grad_b, C, chunks);
    B200_CUDA_LAUNCH_CHECK("gdl_affine_bwd final");
    return B200_OK;
  }
  if (grad_x) {
    // grad_x (in_dtype,in_layout) = grad_y (out_dtype,out_layout) * w * lambda
    int rc = dispatch_affine(grad_y, weight, nullptr, lambda, grad_x, N, C, H, W, out_dtype, out_layout, in_dtype,
                             in_layout, st);
    if (rc != B200_OK) return rc;
  }
  if (grad_w || grad_b) {
    B200_CHECK_ARG(x, "gdl_affine_bwd: x required for parameter gradients");
    if (workspace_bytes < b200_gdl_affine_bwd_workspace_bytes(N, C, H, W) || !workspace) {
      set_error("gdl_affine_bwd: workspace too small");
      return B200_ERR_WORKSPACE;
    }
    float* partial = (float*)workspace;
    dim3 grid(ceil_div(C, 32), kRedChunks), block(32, 8);
    const int HW = H * W;
    if (in_dtype == B200_F32 && out_dtype == B200_F32)
      affine_param_grad_partial_kernel<float, float><<<grid, block, 0, st>>>((const float*)grad_y, (const float*)x, partial, N, C, HW, in_layout, out_layout);
    else if (in_dtype == B200_F32 && out_dtype == B200_BF16)
      affine_param_grad_partial_kernel<float, __nv_bfloat16><<<grid, block, 0, st>>>((const __nv_bfloat16*)grad_y, (const float*)x, partial, N, C, HW, in_layout, out_layout);
    else if (in_dtype == B200_BF16 && out_dtype == B200_F32)
      affine_param_grad_partial_kernel<__nv_bfloat16, float><<<grid, block, 0, st>>>((const float*)grad_y, (const __nv_bfloat16*)x, partial, N, C, HW, in_layout, out_layout);
    else
      affine_param_grad_partial_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, block, 0, st>>>((const __nv_bfloat16*)grad_y, (const __nv_bfloat16*)x, partial, N, C, HW, in_layout, out_layout);
    B200_CUDA_LAUNCH_CHECK("gdl_affine_bwd partial");
    affine_param_grad_final_kernel<<<ceil_div(C, 128), 128, 0, st>>>(partial, grad_w, grad_b, C, kRedChunks);
    B200_CUDA_LAUNCH_CHECK("gdl_affine_bwd final");
  }
  return B200_OK;
}
