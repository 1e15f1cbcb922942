// Fine-tuning direction of the text-fused head: the non-GEMM kernels of the backward pass and the losses.
// Reference (autograd of): defrcn/modeling/roi_heads/attentive_modules.py:45-55,114-177 (attention core, gating,
// FFN + LayerNorm), fast_rcnn.py:222-304 (softmax CE, smooth-L1 on foreground rows / R), roi_heads.py:1077-1081
// (CE over the attention *probabilities*), fast_rcnn.py:412-414 (classifier dropout), and the SGD+momentum update
// of defrcn/solver/build.py (torch.optim.SGD semantics).
// Everything is deterministic: column sums are two-pass, losses are reduced in a fixed order, dropout is a
// counter-based hash of (seed, element index) that the backward pass re-evaluates instead of storing a mask.
#include "common.cuh"

namespace b200 {

__device__ __forceinline__ float t_bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float t_bf16_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
__device__ __forceinline__ uint32_t t_pack(float a, float b) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

// ------------------------------------------------------------------------------------------------
// transpose (+cast): dst[c][r] = bf16(src[r][c]); 64 x 64 tiles through shared memory
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
transpose_bf16_kernel(const T* __restrict__ src, int ld_src, __nv_bfloat16* __restrict__ dst, int ld_dst, int rows,
                      int cols, int vec_ok) {
  __shared__ __nv_bfloat16 tile[64][72];     // [src row][src col]; 144-byte pitch keeps the 16-byte rows aligned
  const int r0 = blockIdx.y * 64, c0 = blockIdx.x * 64;
  // load: 64 rows x 8 groups of 8 columns; thread t -> row t / 8 (+32), column group t % 8
  for (int i = threadIdx.x; i < 64 * 8; i += 256) {
    const int lr = i >> 3, gc = (i & 7) * 8;
    const int r = r0 + lr, c = c0 + gc;
    __nv_bfloat16 v[8];
    if (vec_ok && r < rows && c + 8 <= cols) {
      if (sizeof(T) == 4) {
        const float4 a = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(src) + (size_t)r * ld_src + c);
        const float4 b = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(src) + (size_t)r * ld_src + c + 4);
        v[0] = __float2bfloat16_rn(a.x); v[1] = __float2bfloat16_rn(a.y); v[2] = __float2bfloat16_rn(a.z); v[3] = __float2bfloat16_rn(a.w);
        v[4] = __float2bfloat16_rn(b.x); v[5] = __float2bfloat16_rn(b.y); v[6] = __float2bfloat16_rn(b.z); v[7] = __float2bfloat16_rn(b.w);
      } else {
        *reinterpret_cast<uint4*>(v) = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(src) + (size_t)r * ld_src + c);
      }
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float f = 0.f;
        if (r < rows && c + k < cols) {
          if (sizeof(T) == 4) f = reinterpret_cast<const float*>(src)[(size_t)r * ld_src + c + k];
          else f = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(src)[(size_t)r * ld_src + c + k]);
        }
        v[k] = __float2bfloat16_rn(f);
      }
    }
    // 8-column groups are XOR-swizzled with the row's 8-row block so that the transposed reads below (8 lanes, rows
    // 8 apart, same column) fall on 8 different 16-byte bank groups
    *reinterpret_cast<uint4*>(&tile[lr][gc ^ ((lr >> 3) << 3)]) = *reinterpret_cast<const uint4*>(v);
  }
  __syncthreads();
  // store: destination row = source column; 8 consecutive source rows per 16-byte store
  for (int i = threadIdx.x; i < 64 * 8; i += 256) {
    const int lc = i >> 3, gr = (i & 7) * 8;
    const int c = c0 + lc, r = r0 + gr;
    if (c >= cols) continue;
    __nv_bfloat16 v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = tile[gr + k][lc ^ gr];      // (gr + k) >> 3 == gr >> 3; gr is a multiple of 8
    __nv_bfloat16* d = dst + (size_t)c * ld_dst + r;
    if (r + 8 <= rows && ((reinterpret_cast<uintptr_t>(d) & 15) == 0)) {
      *reinterpret_cast<uint4*>(d) = *reinterpret_cast<const uint4*>(v);
    } else {
      for (int k = 0; k < 8; ++k)
        if (r + k < rows) d[k] = v[k];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// column sums (bias gradients): pass 1 = fixed row blocks -> partial[nblk][cols], pass 2 = ordered sum of the partials
// ------------------------------------------------------------------------------------------------
constexpr int kColBlocks = 64;

template <typename T>
__global__ void __launch_bounds__(256)
colsum_partial_kernel(const T* __restrict__ src, int ld, int rows, int cols, float* __restrict__ partial) {
  const int c = blockIdx.x * 256 + threadIdx.x;
  const int rb = blockIdx.y;
  const int per = (rows + kColBlocks - 1) / kColBlocks;
  const int r0 = rb * per, r1 = min(rows, r0 + per);
  if (c >= cols) return;
  float s = 0.f;
  for (int r = r0; r < r1; ++r) {
    if (sizeof(T) == 4) s += reinterpret_cast<const float*>(src)[(size_t)r * ld + c];
    else s += __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(src)[(size_t)r * ld + c]);
  }
  partial[(size_t)rb * cols + c] = s;
}

__global__ void colsum_final_kernel(const float* __restrict__ partial, int nblk, int cols, float* __restrict__ out,
                                    int accumulate) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  float s = 0.f;
  for (int b = 0; b < nblk; ++b) s += partial[(size_t)b * cols + c];
  out[c] = accumulate ? out[c] + s : s;
}

// ------------------------------------------------------------------------------------------------
// dropout: keep iff hash(seed, index) >= p * 2^32; kept values scaled by 1/(1-p).  Counter based: the backward
// pass regenerates the decision from (seed, index).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t hash32(uint64_t seed, uint64_t idx) {
  uint64_t z = seed + idx * 0x9E3779B97F4A7C15ull + 0x632BE59BD9B4E019ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return (uint32_t)(z >> 32);
}
__device__ __forceinline__ bool dropout_keep(uint64_t seed, uint64_t idx, uint32_t thresh) { return hash32(seed, idx) >= thresh; }
__host__ __device__ inline uint32_t dropout_thresh(float p) {
  const double t = (double)p * 4294967296.0;
  return t >= 4294967295.0 ? 0xffffffffu : (uint32_t)t;
}

__global__ void dropout_fwd_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, size_t n, float p,
                                   unsigned long long seed, const unsigned long long* __restrict__ salt) {
  if (salt) seed += *salt;              // device-resident step counter: a captured CUDA graph draws a new mask per replay
  const uint32_t th = dropout_thresh(p);
  const float scale = p < 1.f ? 1.f / (1.f - p) : 0.f;
  for (size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 2; i < n; i += (size_t)gridDim.x * blockDim.x * 2) {
    const float a = (p > 0.f && !dropout_keep(seed, i, th)) ? 0.f : x[i] * (p > 0.f ? scale : 1.f);
    float b = 0.f;
    if (i + 1 < n) b = (p > 0.f && !dropout_keep(seed, i + 1, th)) ? 0.f : x[i + 1] * (p > 0.f ? scale : 1.f);
    if (i + 1 < n) *reinterpret_cast<uint32_t*>(y + i) = t_pack(a, b);
    else y[i] = __float2bfloat16_rn(a);
  }
}

// zd = dropout(relu(LayerNorm(y + y2) * gamma + beta)) as bf16 in one pass (attentive_modules.py:73-74,285 followed by the
// classifier dropout of fast_rcnn.py:412-414): the fp32 z of the two-kernel form is never written or re-read (66 MB per step at
// R = 4096).  Statistics, affine map and dropout decision are those of residual_layernorm_kernel / dropout_fwd_kernel,
// expression for expression, so the result equals the two-kernel form bit for bit and the backward's recomputation matches.
__global__ void __launch_bounds__(256)
residual_layernorm_dropout_kernel(const float* __restrict__ y, const float* __restrict__ y2, const float* __restrict__ gamma,
                                  const float* __restrict__ beta, float eps, int relu, float p, unsigned long long seed,
                                  const unsigned long long* __restrict__ salt, __nv_bfloat16* __restrict__ out, int R, int d) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= R) return;
  if (salt) seed += *salt;
  const uint32_t th = dropout_thresh(p);
  const float scale = p < 1.f ? 1.f / (1.f - p) : 0.f;
  const float4* a = reinterpret_cast<const float4*>(y + (size_t)row * d);
  const float4* b = y2 ? reinterpret_cast<const float4*>(y2 + (size_t)row * d) : nullptr;
  const int n4 = d / 4;
  float s = 0.f;
  for (int i = lane; i < n4; i += 32) {
    float4 v = a[i];
    if (b) { const float4 w = b[i]; v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w; }
    s += (v.x + v.y) + (v.z + v.w);
  }
  const float mean = warp_sum(s) / (float)d;
  float q = 0.f;
  for (int i = lane; i < n4; i += 32) {
    float4 v = a[i];
    if (b) { const float4 w = b[i]; v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w; }
    const float dx = v.x - mean, dy = v.y - mean, dz = v.z - mean, dw = v.w - mean;
    q += (dx * dx + dy * dy) + (dz * dz + dw * dw);
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)d + eps);
  for (int i = lane; i < n4; i += 32) {
    float4 v = a[i];
    if (b) { const float4 w = b[i]; v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w; }
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + i);
    const float4 be = __ldg(reinterpret_cast<const float4*>(beta) + i);
    float o[4];
    o[0] = (v.x - mean) * rstd * g.x + be.x; o[1] = (v.y - mean) * rstd * g.y + be.y;
    o[2] = (v.z - mean) * rstd * g.z + be.z; o[3] = (v.w - mean) * rstd * g.w + be.w;
    const size_t idx = (size_t)row * d + 4 * (size_t)i;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (relu) o[k] = fmaxf(o[k], 0.f);
      o[k] = (p > 0.f && !dropout_keep(seed, idx + k, th)) ? 0.f : o[k] * (p > 0.f ? scale : 1.f);
    }
    uint2 w;
    w.x = t_pack(o[0], o[1]); w.y = t_pack(o[2], o[3]);
    reinterpret_cast<uint2*>(out + (size_t)row * d)[i] = w;
  }
}

// ------------------------------------------------------------------------------------------------
// backward of  z = relu(LayerNorm(y + y2) * gamma + beta), zd = dropout(z):
//   dz = dzd * keep/(1-p) * (z > 0);  du = rstd * (dxh - mean(dxh) - xh * mean(dxh * xh)),  dxh = dz * gamma
// kernel 1: one warp per row -> du (fp32 and/or bf16) and the row statistics (mean, rstd);
// kernel 2: dgamma = sum_r dz * xh, dbeta = sum_r dz as fixed row-block partials (thread per column), summed in order
//           by colsum_final_kernel.  dz is re-derived from dzd / the dropout hash / sign(z) instead of being stored.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float ln_dz(const __nv_bfloat16* dzd, size_t idx, float zv, float p, unsigned long long seed,
                                       uint32_t th, float dscale) {
  float dz = __bfloat162float(dzd[idx]);
  if (p > 0.f) dz = dropout_keep(seed, idx, th) ? dz * dscale : 0.f;
  return zv > 0.f ? dz : 0.f;
}

__global__ void __launch_bounds__(256)
layernorm_relu_dropout_bwd_kernel(const __nv_bfloat16* __restrict__ dzd, const float* __restrict__ y,
                                  const float* __restrict__ y2, const float* __restrict__ gamma,
                                  const float* __restrict__ beta, float eps, float p, unsigned long long seed,
                                  const unsigned long long* __restrict__ salt, float* __restrict__ du_f32,
                                  __nv_bfloat16* __restrict__ du_bf16, float* __restrict__ stats, int R, int d) {
  if (salt) seed += *salt;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= R) return;
  const uint32_t th = dropout_thresh(p);
  const float dscale = p > 0.f ? (p < 1.f ? 1.f / (1.f - p) : 0.f) : 1.f;
  const float* a = y + (size_t)row * d;
  const float* b = y2 + (size_t)row * d;
  float s = 0.f;
  for (int i = lane; i < d; i += 32) s += a[i] + b[i];
  const float mean = warp_sum(s) / (float)d;
  float q = 0.f;
  for (int i = lane; i < d; i += 32) { const float dx = a[i] + b[i] - mean; q += dx * dx; }
  const float rstd = rsqrtf(warp_sum(q) / (float)d + eps);
  if (lane == 0) { stats[2 * row] = mean; stats[2 * row + 1] = rstd; }
  float m1 = 0.f, m2 = 0.f;
  for (int i = lane; i < d; i += 32) {
    const float xh = (a[i] + b[i] - mean) * rstd;
    const float dxh = ln_dz(dzd, (size_t)row * d + i, xh * gamma[i] + beta[i], p, seed, th, dscale) * gamma[i];
    m1 += dxh; m2 += dxh * xh;
  }
  m1 = warp_sum(m1) / (float)d; m2 = warp_sum(m2) / (float)d;
  for (int i = lane; i < d; i += 32) {
    const float xh = (a[i] + b[i] - mean) * rstd;
    const float dxh = ln_dz(dzd, (size_t)row * d + i, xh * gamma[i] + beta[i], p, seed, th, dscale) * gamma[i];
    const float du = rstd * (dxh - m1 - xh * m2);
    if (du_f32) du_f32[(size_t)row * d + i] = du;
    if (du_bf16) du_bf16[(size_t)row * d + i] = __float2bfloat16_rn(du);
  }
}

// Register-resident variant for d = 128 V: a lane holds V float4 of its row, so y, y2 and dzd are read once (16-byte / 8-byte
// accesses, all of a row's loads in flight together) instead of four strided passes over the row.
template <int V>
__global__ void __launch_bounds__(256)
layernorm_relu_dropout_bwd_vec_kernel(const __nv_bfloat16* __restrict__ dzd, const float* __restrict__ y,
                                      const float* __restrict__ y2, const float* __restrict__ gamma,
                                      const float* __restrict__ beta, float eps, float p, unsigned long long seed,
                                      const unsigned long long* __restrict__ salt, float* __restrict__ du_f32,
                                      __nv_bfloat16* __restrict__ du_bf16, float* __restrict__ stats, int R) {
  constexpr int d = 128 * V;
  if (salt) seed += *salt;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= R) return;
  const uint32_t th = dropout_thresh(p);
  const float dscale = p > 0.f ? (p < 1.f ? 1.f / (1.f - p) : 0.f) : 1.f;
  const float4* a = reinterpret_cast<const float4*>(y + (size_t)row * d);
  const float4* b = reinterpret_cast<const float4*>(y2 + (size_t)row * d);
  const uint2* gz = reinterpret_cast<const uint2*>(dzd + (size_t)row * d);
  float4 u[V];
  uint2 z[V];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const float4 pa = __ldcs(a + i * 32 + lane), pb = __ldcs(b + i * 32 + lane);
    z[i] = __ldcs(gz + i * 32 + lane);
    u[i] = make_float4(pa.x + pb.x, pa.y + pb.y, pa.z + pb.z, pa.w + pb.w);
    s += (u[i].x + u[i].y) + (u[i].z + u[i].w);
  }
  const float mean = warp_sum(s) / (float)d;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const float dx = u[i].x - mean, dy = u[i].y - mean, dz = u[i].z - mean, dw = u[i].w - mean;
    q += (dx * dx + dy * dy) + (dz * dz + dw * dw);
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)d + eps);
  if (lane == 0) { stats[2 * row] = mean; stats[2 * row + 1] = rstd; }
  float m1 = 0.f, m2 = 0.f;
  float4 dxh[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const int col = (i * 32 + lane) * 4;
    const float4 g4 = __ldg(reinterpret_cast<const float4*>(gamma + col)), b4 = __ldg(reinterpret_cast<const float4*>(beta + col));
    const float gv[4] = {g4.x, g4.y, g4.z, g4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
    float xv[4] = {u[i].x, u[i].y, u[i].z, u[i].w};
    const float zv[4] = {__uint_as_float(z[i].x << 16), __uint_as_float(z[i].x & 0xffff0000u), __uint_as_float(z[i].y << 16),
                         __uint_as_float(z[i].y & 0xffff0000u)};
    float dv[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float xh = (xv[k] - mean) * rstd;
      float dz = zv[k];
      if (p > 0.f) dz = dropout_keep(seed, (size_t)row * d + col + k, th) ? dz * dscale : 0.f;
      if (!(xh * gv[k] + bv[k] > 0.f)) dz = 0.f;
      dv[k] = dz * gv[k];
      m1 += dv[k];
      m2 += dv[k] * xh;
      xv[k] = xh;
    }
    u[i] = make_float4(xv[0], xv[1], xv[2], xv[3]);
    dxh[i] = make_float4(dv[0], dv[1], dv[2], dv[3]);
  }
  m1 = warp_sum(m1) / (float)d;
  m2 = warp_sum(m2) / (float)d;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const int col = (i * 32 + lane) * 4;
    float4 o;
    o.x = rstd * (dxh[i].x - m1 - u[i].x * m2); o.y = rstd * (dxh[i].y - m1 - u[i].y * m2);
    o.z = rstd * (dxh[i].z - m1 - u[i].z * m2); o.w = rstd * (dxh[i].w - m1 - u[i].w * m2);
    if (du_f32) *reinterpret_cast<float4*>(du_f32 + (size_t)row * d + col) = o;
    if (du_bf16) {
      uint2 w;
      w.x = t_pack(o.x, o.y); w.y = t_pack(o.z, o.w);
      *reinterpret_cast<uint2*>(du_bf16 + (size_t)row * d + col) = w;
    }
  }
}

__global__ void __launch_bounds__(256)
layernorm_param_grad_partial_kernel(const __nv_bfloat16* __restrict__ dzd, const float* __restrict__ y,
                                    const float* __restrict__ y2, const float* __restrict__ gamma,
                                    const float* __restrict__ beta, const float* __restrict__ stats, float p,
                                    unsigned long long seed, const unsigned long long* __restrict__ salt,
                                    float* __restrict__ part_g, float* __restrict__ part_b, int R, int d) {
  if (salt) seed += *salt;
  const int c = blockIdx.x * 256 + threadIdx.x;
  const int rb = blockIdx.y;
  const int per = (R + kColBlocks - 1) / kColBlocks;
  const int r0 = rb * per, r1 = min(R, r0 + per);
  if (c >= d) return;
  const uint32_t th = dropout_thresh(p);
  const float dscale = p > 0.f ? (p < 1.f ? 1.f / (1.f - p) : 0.f) : 1.f;
  const float g = gamma[c], be = beta[c];
  float sg = 0.f, sb = 0.f;
  for (int r = r0; r < r1; ++r) {
    const size_t idx = (size_t)r * d + c;
    const float xh = (y[idx] + y2[idx] - stats[2 * r]) * stats[2 * r + 1];
    const float dz = ln_dz(dzd, idx, xh * g + be, p, seed, th, dscale);
    sg += dz * xh; sb += dz;
  }
  part_g[(size_t)rb * d + c] = sg;
  part_b[(size_t)rb * d + c] = sb;
}

// ------------------------------------------------------------------------------------------------
// backward of the attention core + gate operands (forward: text_attention_kernel in fusion_elem.cu):
//   O = attn Vp,  P1 = O * x,  P2 = x - O
//   dO = dP1 * x - dP2;  dx (+)= dP1 * O + dP2;  dA = dO Vp^T;  g = dA + dattn_ext
//   dS = attn * (g - sum_l attn_l g_l)                                  (softmax backward)
// CTA = TR rows; thread owns 8 columns of all TR rows (O is recomputed, never stored by the forward).
// dVp = attn^T dO is a [L x R][R x d] contraction and goes through the GEMM on the transposed copies.
// ------------------------------------------------------------------------------------------------
constexpr int kAbThreads = 256;
constexpr int kAbRows = 8;
constexpr int kAbMaxL = 128;

__global__ void __launch_bounds__(kAbThreads)
text_attention_bwd_kernel(const __nv_bfloat16* __restrict__ dp1, const __nv_bfloat16* __restrict__ dp2, int ldp,
                          const float* __restrict__ x, const float* __restrict__ attn, const float* __restrict__ vp,
                          const float* __restrict__ dattn_ext, float* __restrict__ dx, int accumulate_dx,
                          __nv_bfloat16* __restrict__ d_o, __nv_bfloat16* __restrict__ ds, int ldds, int R, int d, int L) {
  __shared__ float s_attn[kAbRows][kAbMaxL];
  __shared__ float s_part[kAbThreads / 32][kAbRows][kAbMaxL];
  const int r0 = blockIdx.x * kAbRows;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < kAbRows * L; i += kAbThreads) {
    const int row = i / L, l = i - row * L;
    s_attn[row][l] = (r0 + row < R) ? attn[(size_t)(r0 + row) * L + l] : 0.f;
  }
  for (int i = threadIdx.x; i < (kAbThreads / 32) * kAbRows * kAbMaxL; i += kAbThreads) (&s_part[0][0][0])[i] = 0.f;
  __syncthreads();
  for (int cb = 0; cb < d; cb += kAbThreads * 8) {     // warp-uniform trip count: the reductions below are warp-wide
    const int c = min(cb + (int)threadIdx.x * 8, d - 8);
    const bool active = cb + (int)threadIdx.x * 8 < d;
    float o[kAbRows][8];
#pragma unroll
    for (int row = 0; row < kAbRows; ++row)
#pragma unroll
      for (int k = 0; k < 8; ++k) o[row][k] = 0.f;
    for (int l = 0; l < L; ++l) {
      const float4 va = __ldg(reinterpret_cast<const float4*>(vp + (size_t)l * d + c));
      const float4 vb = __ldg(reinterpret_cast<const float4*>(vp + (size_t)l * d + c + 4));
      const float vv[8] = {va.x, va.y, va.z, va.w, vb.x, vb.y, vb.z, vb.w};
#pragma unroll
      for (int row = 0; row < kAbRows; ++row) {
        const float a = s_attn[row][l];
#pragma unroll
        for (int k = 0; k < 8; ++k) o[row][k] += a * vv[k];
      }
    }
    // o -> dO in place; dx and dO leave here
#pragma unroll
    for (int row = 0; row < kAbRows; ++row) {
      const int r = r0 + row;
      if (r >= R || !active) {
#pragma unroll
        for (int k = 0; k < 8; ++k) o[row][k] = 0.f;
        continue;
      }
      const uint4 t1 = *reinterpret_cast<const uint4*>(dp1 + (size_t)r * ldp + c);
      const uint4 t2 = *reinterpret_cast<const uint4*>(dp2 + (size_t)r * ldp + c);
      const float g1[8] = {t_bf16_lo(t1.x), t_bf16_hi(t1.x), t_bf16_lo(t1.y), t_bf16_hi(t1.y),
                           t_bf16_lo(t1.z), t_bf16_hi(t1.z), t_bf16_lo(t1.w), t_bf16_hi(t1.w)};
      const float g2[8] = {t_bf16_lo(t2.x), t_bf16_hi(t2.x), t_bf16_lo(t2.y), t_bf16_hi(t2.y),
                           t_bf16_lo(t2.z), t_bf16_hi(t2.z), t_bf16_lo(t2.w), t_bf16_hi(t2.w)};
      const float4 xa = *reinterpret_cast<const float4*>(x + (size_t)r * d + c);
      const float4 xb = *reinterpret_cast<const float4*>(x + (size_t)r * d + c + 4);
      const float xx[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
      float dxv[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        dxv[k] = g1[k] * o[row][k] + g2[k];
        o[row][k] = g1[k] * xx[k] - g2[k];
      }
      float* dxp = dx + (size_t)r * d + c;
      if (accumulate_dx) {
        const float4 pa = *reinterpret_cast<const float4*>(dxp), pb = *reinterpret_cast<const float4*>(dxp + 4);
        dxv[0] += pa.x; dxv[1] += pa.y; dxv[2] += pa.z; dxv[3] += pa.w;
        dxv[4] += pb.x; dxv[5] += pb.y; dxv[6] += pb.z; dxv[7] += pb.w;
      }
      *reinterpret_cast<float4*>(dxp) = make_float4(dxv[0], dxv[1], dxv[2], dxv[3]);
      *reinterpret_cast<float4*>(dxp + 4) = make_float4(dxv[4], dxv[5], dxv[6], dxv[7]);
      uint4 w;
      w.x = t_pack(o[row][0], o[row][1]); w.y = t_pack(o[row][2], o[row][3]);
      w.z = t_pack(o[row][4], o[row][5]); w.w = t_pack(o[row][6], o[row][7]);
      *reinterpret_cast<uint4*>(d_o + (size_t)r * d + c) = w;
    }
    // dA partials: this thread's 8 columns, reduced over the warp, accumulated per warp (fixed order)
    for (int l = 0; l < L; ++l) {
      const float4 va = __ldg(reinterpret_cast<const float4*>(vp + (size_t)l * d + c));
      const float4 vb = __ldg(reinterpret_cast<const float4*>(vp + (size_t)l * d + c + 4));
      const float vv[8] = {va.x, va.y, va.z, va.w, vb.x, vb.y, vb.z, vb.w};
#pragma unroll
      for (int row = 0; row < kAbRows; ++row) {
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) s += o[row][k] * vv[k];
        s = warp_sum(s);
        if (lane == 0) s_part[warp][row][l] += s;
      }
    }
  }
  __syncthreads();
  // softmax backward: warp <-> row
  for (int row = warp; row < kAbRows; row += kAbThreads / 32) {
    const int r = r0 + row;
    if (r >= R) continue;
    float g[kAbMaxL / 32], a[kAbMaxL / 32];
    float dot = 0.f;
#pragma unroll
    for (int j = 0; j < kAbMaxL / 32; ++j) {
      const int l = lane + 32 * j;
      g[j] = 0.f; a[j] = 0.f;
      if (l < L) {
        float t = 0.f;
        for (int w = 0; w < kAbThreads / 32; ++w) t += s_part[w][row][l];
        if (dattn_ext) t += dattn_ext[(size_t)r * L + l];
        g[j] = t; a[j] = s_attn[row][l];
        dot += a[j] * t;
      }
    }
    dot = warp_sum(dot);
#pragma unroll
    for (int j = 0; j < kAbMaxL / 32; ++j) {
      const int l = lane + 32 * j;
      if (l < ldds) ds[(size_t)r * ldds + l] = __float2bfloat16_rn(l < L ? a[j] * (g[j] - dot) : 0.f);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// losses (fast_rcnn.py:222-304, roi_heads.py:1079-1081): one CTA, rows strided over the threads, fixed-order tree
//   out[0] = mean_r CE(logits_r, gt_r)
//   out[1] = sum_{r fg} sum_{j<4} smooth_l1(deltas[r, 4 gt_r + j] - target[r, j]) / R      (beta = 0: |.|)
//   out[2] = mean_r CE(attn_r, gt_r)              — CE over the attention PROBABILITIES, as the reference does
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void box_deltas(const float* src, const float* dst, float wx, float wy, float ww, float wh,
                                           float* t) {
  const float sw = src[2] - src[0], sh = src[3] - src[1];
  const float sx = src[0] + 0.5f * sw, sy = src[1] + 0.5f * sh;
  const float tw = dst[2] - dst[0], thh = dst[3] - dst[1];
  const float tx = dst[0] + 0.5f * tw, ty = dst[1] + 0.5f * thh;
  t[0] = wx * (tx - sx) / sw; t[1] = wy * (ty - sy) / sh; t[2] = ww * logf(tw / sw); t[3] = wh * logf(thh / sh);
}
__device__ __forceinline__ float row_lse(const float* v, int n) {
  float mx = -INFINITY;
  for (int i = 0; i < n; ++i) mx = fmaxf(mx, v[i]);
  float s = 0.f;
  for (int i = 0; i < n; ++i) s += expf(v[i] - mx);
  return mx + logf(s);
}

__global__ void __launch_bounds__(1024)
head_losses_kernel(const float* __restrict__ logits, const float* __restrict__ deltas, const float* __restrict__ attn,
                   const int64_t* __restrict__ gt, const float* __restrict__ props, const float* __restrict__ gt_boxes,
                   int R, int K, int L, int agnostic, float wx, float wy, float ww, float wh, float beta,
                   float* __restrict__ out, float* __restrict__ acc_stats) {
  __shared__ float s_red[3][1024];
  __shared__ int s_acc[4];
  float lc = 0.f, lb = 0.f, la = 0.f;
  int n_hit = 0, n_fg = 0, n_fg_hit = 0, n_fg_bg = 0;
  const int C1 = K + 1, C4 = agnostic ? 4 : 4 * K;
  if (threadIdx.x < 4) s_acc[threadIdx.x] = 0;
  for (int r = threadIdx.x; r < R; r += blockDim.x) {
    const int g = (int)gt[r];
    lc += row_lse(logits + (size_t)r * C1, C1) - logits[(size_t)r * C1 + g];
    if (acc_stats) {        // FastRCNNOutputs._log_accuracy (fast_rcnn.py:191-220): argmax = first maximum, like torch
      const float* v = logits + (size_t)r * C1;
      int am = 0;
      float best = v[0];
      for (int c = 1; c < C1; ++c) if (v[c] > best) { best = v[c]; am = c; }
      const bool fg = g >= 0 && g < K;
      n_hit += am == g;
      n_fg += fg;
      n_fg_hit += fg && am == g;
      n_fg_bg += fg && am == K;
    }
    if (attn) la += row_lse(attn + (size_t)r * L, L) - attn[(size_t)r * L + g];
    if (g >= 0 && g < K) {
      float t[4];
      box_deltas(props + 4 * (size_t)r, gt_boxes + 4 * (size_t)r, wx, wy, ww, wh, t);
      const float* dp = deltas + (size_t)r * C4 + (agnostic ? 0 : 4 * g);
      for (int j = 0; j < 4; ++j) {
        const float n = fabsf(dp[j] - t[j]);
        lb += beta < 1e-5f ? n : (n < beta ? 0.5f * n * n / beta : n - 0.5f * beta);
      }
    }
  }
  s_red[0][threadIdx.x] = lc; s_red[1][threadIdx.x] = lb; s_red[2][threadIdx.x] = la;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) {
      s_red[0][threadIdx.x] += s_red[0][threadIdx.x + o];
      s_red[1][threadIdx.x] += s_red[1][threadIdx.x + o];
      s_red[2][threadIdx.x] += s_red[2][threadIdx.x + o];
    }
    __syncthreads();
  }
  if (acc_stats) {          // integer counts: the order of the atomic adds does not matter
    __syncthreads();
    if (n_hit) atomicAdd(&s_acc[0], n_hit);
    if (n_fg) atomicAdd(&s_acc[1], n_fg);
    if (n_fg_hit) atomicAdd(&s_acc[2], n_fg_hit);
    if (n_fg_bg) atomicAdd(&s_acc[3], n_fg_bg);
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    out[0] = s_red[0][0] / (float)R; out[1] = s_red[1][0] / (float)R; out[2] = s_red[2][0] / (float)R;
    if (acc_stats) {
      acc_stats[0] = (float)s_acc[0]; acc_stats[1] = (float)s_acc[1]; acc_stats[2] = (float)s_acc[2];
      acc_stats[3] = (float)s_acc[3]; acc_stats[4] = (float)R;
    }
  }
}

// gradients of the three losses w.r.t. logits / deltas / attention probabilities, scaled by the upstream scalars
// gscale[0..2] (device memory: no host sync).  dlogits (R, ldl) and ddeltas (R, ldd) are bf16, zero padded to ld.
__global__ void __launch_bounds__(256)
head_losses_bwd_kernel(const float* __restrict__ logits, const float* __restrict__ deltas, const float* __restrict__ attn,
                       const int64_t* __restrict__ gt, const float* __restrict__ props, const float* __restrict__ gt_boxes,
                       const float* __restrict__ gscale, int R, int K, int L, int agnostic, float wx, float wy, float ww,
                       float wh, float beta, __nv_bfloat16* __restrict__ dlogits, int ldl,
                       __nv_bfloat16* __restrict__ ddeltas, int ldd, float* __restrict__ dattn) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const int C1 = K + 1, C4 = agnostic ? 4 : 4 * K;
  const int g = (int)gt[r];
  const float invR = 1.0f / (float)R;
  {
    const float* v = logits + (size_t)r * C1;
    const float lse = row_lse(v, C1);
    const float s = gscale[0] * invR;
    for (int c = 0; c < ldl; ++c)
      dlogits[(size_t)r * ldl + c] = __float2bfloat16_rn(c < C1 ? s * (expf(v[c] - lse) - (c == g ? 1.f : 0.f)) : 0.f);
  }
  if (attn && dattn) {
    const float* v = attn + (size_t)r * L;
    const float lse = row_lse(v, L);
    const float s = gscale[2] * invR;
    for (int c = 0; c < L; ++c) dattn[(size_t)r * L + c] = s * (expf(v[c] - lse) - (c == g ? 1.f : 0.f));
  }
  {
    float t[4] = {0.f, 0.f, 0.f, 0.f};
    const bool fg = g >= 0 && g < K;
    if (fg) box_deltas(props + 4 * (size_t)r, gt_boxes + 4 * (size_t)r, wx, wy, ww, wh, t);
    const int c0 = agnostic ? 0 : 4 * g;
    const float s = gscale[1] * invR;
    for (int c = 0; c < ldd; ++c) {
      float gv = 0.f;
      if (fg && c >= c0 && c < c0 + 4 && c < C4) {
        const float df = deltas[(size_t)r * C4 + c] - t[c - c0];
        const float n = fabsf(df);
        const float sgn = df > 0.f ? 1.f : (df < 0.f ? -1.f : 0.f);
        gv = s * (beta < 1e-5f ? sgn : (n < beta ? df / beta : sgn));
      }
      ddeltas[(size_t)r * ldd + c] = __float2bfloat16_rn(gv);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// distillation loss of the student head (my_module.py:409-437 loss_fn_kd_only, called from roi_heads.py:760 with
// alpha = 1 and T = KL_TEMP): out[0] = alpha T^2 / R * sum_r w_r sum_c pt_rc (log pt_rc - log ps_rc), with
// pt = softmax(teacher / T), ps = softmax(student / T), w_r = 1.5 on background rows (gt_r == bg_label), else 1.
// One CTA, rows strided over the threads, fixed-order tree (bitwise reproducible).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
kd_loss_kernel(const float* __restrict__ student, const float* __restrict__ teacher, const int64_t* __restrict__ gt, int R,
               int C1, int bg_label, float T, float alpha, float* __restrict__ out) {
  __shared__ float s_red[1024];
  const float invT = 1.0f / T;
  float acc = 0.f;
  for (int r = threadIdx.x; r < R; r += blockDim.x) {
    const float* s = student + (size_t)r * C1;
    const float* t = teacher + (size_t)r * C1;
    float ms = -INFINITY, mt = -INFINITY;
    for (int c = 0; c < C1; ++c) { ms = fmaxf(ms, s[c] * invT); mt = fmaxf(mt, t[c] * invT); }
    float zs = 0.f, zt = 0.f;
    for (int c = 0; c < C1; ++c) { zs += expf(s[c] * invT - ms); zt += expf(t[c] * invT - mt); }
    const float ls = ms + logf(zs), lt = mt + logf(zt);
    float kl = 0.f;
    for (int c = 0; c < C1; ++c) {
      const float lpt = t[c] * invT - lt;
      const float pt = expf(lpt);
      if (pt > 0.f) kl += pt * (lpt - (s[c] * invT - ls));
    }
    acc += ((int)gt[r] == bg_label) ? 1.5f * kl : kl;
  }
  s_red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) s_red[threadIdx.x] += s_red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = s_red[0] / (float)R * T * T * alpha;
}

// dlogits[r][c] (bf16, the student's logit gradient left by head_losses_bwd) += gscale[0] w_r alpha T / R (ps_rc - pt_rc)
__global__ void __launch_bounds__(256)
kd_loss_bwd_kernel(const float* __restrict__ student, const float* __restrict__ teacher, const int64_t* __restrict__ gt,
                   const float* __restrict__ gscale, int R, int C1, int bg_label, float T, float alpha,
                   __nv_bfloat16* __restrict__ dlogits, int ldl) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const float invT = 1.0f / T;
  const float* s = student + (size_t)r * C1;
  const float* t = teacher + (size_t)r * C1;
  float ms = -INFINITY, mt = -INFINITY;
  for (int c = 0; c < C1; ++c) { ms = fmaxf(ms, s[c] * invT); mt = fmaxf(mt, t[c] * invT); }
  float zs = 0.f, zt = 0.f;
  for (int c = 0; c < C1; ++c) { zs += expf(s[c] * invT - ms); zt += expf(t[c] * invT - mt); }
  const float w = (((int)gt[r] == bg_label) ? 1.5f : 1.0f) * gscale[0] * alpha * T / (float)R;
  for (int c = 0; c < C1; ++c) {
    const float ps = expf(s[c] * invT - ms) / zs, pt = expf(t[c] * invT - mt) / zt;
    __nv_bfloat16* d = dlogits + (size_t)r * ldl + c;
    *d = __float2bfloat16_rn(__bfloat162float(*d) + w * (ps - pt));
  }
}

// SGD with momentum over one flat fp32 buffer (torch.optim.SGD: g += wd * p; m = mu * m + g; p -= lr * m)
// `shadow` (optional): the bf16 copy of the updated parameters the tensor-core GEMMs read as operands, written in the
// same pass instead of by one cast kernel per weight matrix at the top of the next step.
__global__ void sgd_momentum_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, size_t n,
                                    float lr, float mu, float wd, __nv_bfloat16* __restrict__ shadow) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float gi = g[i] + wd * p[i];
    const float mi = mu * m[i] + gi;
    m[i] = mi;
    const float pi = p[i] - lr * mi;
    p[i] = pi;
    if (shadow) shadow[i] = __float2bfloat16_rn(pi);
  }
}

}  // namespace b200

using namespace b200;

extern "C" int b200_transpose_bf16(const void* src, int src_dtype, int ld_src, void* dst, int ld_dst, int rows, int cols,
                                   b200_stream_t stream) {
  B200_CHECK_ARG(rows >= 0 && cols >= 0 && (src_dtype | 1) == 1, "transpose_bf16: bad arguments");
  if (rows == 0 || cols == 0) return B200_OK;
  B200_CHECK_ARG(src && dst && ld_src >= cols && ld_dst >= rows, "transpose_bf16: null tensor or short leading dimension");
  dim3 grid(ceil_div(cols, 64), ceil_div(rows, 64));
  const int esz = src_dtype == B200_F32 ? 4 : 2;
  const int vec_ok = (((uintptr_t)src & 15) == 0) && ((size_t)ld_src * esz % 16 == 0);
  if (src_dtype == B200_F32)
    transpose_bf16_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)src, ld_src, (__nv_bfloat16*)dst, ld_dst, rows, cols, vec_ok);
  else
    transpose_bf16_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)src, ld_src, (__nv_bfloat16*)dst, ld_dst, rows, cols, vec_ok);
  B200_CUDA_LAUNCH_CHECK("transpose_bf16");
  return B200_OK;
}

extern "C" size_t b200_colsum_workspace_bytes(int cols) { return (size_t)kColBlocks * (size_t)max(cols, 1) * 4 * 2; }

extern "C" int b200_colsum(const void* src, int src_dtype, int ld, int rows, int cols, float* out, int accumulate,
                           void* workspace, size_t workspace_bytes, b200_stream_t stream) {
  B200_CHECK_ARG(rows >= 0 && cols > 0 && out && (src_dtype | 1) == 1, "colsum: bad arguments");
  B200_CHECK_ARG(workspace && workspace_bytes >= (size_t)kColBlocks * cols * 4, "colsum: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  float* part = (float*)workspace;
  dim3 grid(ceil_div(cols, 256), kColBlocks);
  if (src_dtype == B200_F32) colsum_partial_kernel<float><<<grid, 256, 0, st>>>((const float*)src, ld, rows, cols, part);
  else colsum_partial_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)src, ld, rows, cols, part);
  colsum_final_kernel<<<ceil_div(cols, 256), 256, 0, st>>>(part, kColBlocks, cols, out, accumulate);
  B200_CUDA_LAUNCH_CHECK("colsum");
  return B200_OK;
}

extern "C" int b200_dropout_fwd(const float* x, void* y_bf16, size_t n, float p, unsigned long long seed,
                                const unsigned long long* seed_salt, b200_stream_t stream) {
  B200_CHECK_ARG(p >= 0.f && p <= 1.f, "dropout: p must be in [0, 1]");
  if (n == 0) return B200_OK;
  B200_CHECK_ARG(x && y_bf16, "dropout: null tensor");
  const int blocks = (int)min((size_t)kNumSMs * 8, (n / 2 + 255) / 256 + 1);
  dropout_fwd_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, (__nv_bfloat16*)y_bf16, n, p, seed, seed_salt);
  B200_CUDA_LAUNCH_CHECK("dropout_fwd");
  return B200_OK;
}

extern "C" int b200_residual_layernorm_dropout(const float* y, const float* y2, const float* gamma, const float* beta, float eps,
                                               int relu, float p, unsigned long long seed, const unsigned long long* seed_salt,
                                               void* out_bf16, int R, int d, b200_stream_t stream) {
  B200_CHECK_ARG(y && gamma && beta && out_bf16, "residual_layernorm_dropout: null tensor");
  B200_CHECK_ARG(R >= 0 && d > 0 && d % 4 == 0, "residual_layernorm_dropout: d must be a multiple of 4");
  B200_CHECK_ARG(p >= 0.f && p <= 1.f, "residual_layernorm_dropout: p must be in [0, 1]");
  if ((((uintptr_t)y | (uintptr_t)y2 | (uintptr_t)gamma | (uintptr_t)beta) & 15) || ((uintptr_t)out_bf16 & 7)) {
    set_error("residual_layernorm_dropout: fp32 operands must be 16-byte, the bf16 output 8-byte aligned");
    return B200_ERR_UNSUPPORTED;
  }
  if (R == 0) return B200_OK;
  residual_layernorm_dropout_kernel<<<ceil_div(R, 8), 256, 0, (cudaStream_t)stream>>>(y, y2, gamma, beta, eps, relu, p, seed, seed_salt,
                                                                                      (__nv_bfloat16*)out_bf16, R, d);
  B200_CUDA_LAUNCH_CHECK("residual_layernorm_dropout");
  return B200_OK;
}

extern "C" size_t b200_layernorm_bwd_workspace_bytes(int R, int d) {
  return align_up((size_t)max(R, 1) * 2 * 4, 256) + 2 * (size_t)kColBlocks * d * 4;
}

extern "C" int b200_layernorm_relu_dropout_bwd(const void* dzd_bf16, const float* y, const float* y2, const float* gamma,
                                               const float* beta, float eps, float p, unsigned long long seed,
                                               const unsigned long long* seed_salt, float* du_f32, void* du_bf16,
                                               float* dgamma, float* dbeta, int R, int d, void* workspace,
                                               size_t workspace_bytes, b200_stream_t stream) {
  B200_CHECK_ARG(dzd_bf16 && y && y2 && gamma && beta && (du_f32 || du_bf16), "layernorm_bwd: null tensor");
  B200_CHECK_ARG(R >= 0 && d > 0, "layernorm_bwd: bad shape");
  B200_CHECK_ARG(workspace && workspace_bytes >= b200_layernorm_bwd_workspace_bytes(R, d), "layernorm_bwd: workspace too small");
  if (R == 0) return B200_OK;
  cudaStream_t st = (cudaStream_t)stream;
  float* stats = (float*)workspace;
  float* pg = (float*)((unsigned char*)workspace + align_up((size_t)R * 2 * 4, 256));
  float* pb = pg + (size_t)kColBlocks * d;
  const bool aligned = ((((uintptr_t)dzd_bf16 | (uintptr_t)du_bf16) & 7) | (((uintptr_t)y | (uintptr_t)y2 | (uintptr_t)gamma |
                                                                             (uintptr_t)beta | (uintptr_t)du_f32) & 15)) == 0;
  if (d == 2048 && aligned)
    layernorm_relu_dropout_bwd_vec_kernel<16><<<ceil_div(R, 8), 256, 0, st>>>((const __nv_bfloat16*)dzd_bf16, y, y2, gamma, beta,
                                                                              eps, p, seed, seed_salt, du_f32,
                                                                              (__nv_bfloat16*)du_bf16, stats, R);
  else
    layernorm_relu_dropout_bwd_kernel<<<ceil_div(R, 8), 256, 0, st>>>((const __nv_bfloat16*)dzd_bf16, y, y2, gamma, beta, eps, p,
                                                                      seed, seed_salt, du_f32, (__nv_bfloat16*)du_bf16, stats, R, d);
  if (dgamma || dbeta) {
    dim3 grid(ceil_div(d, 256), kColBlocks);
    layernorm_param_grad_partial_kernel<<<grid, 256, 0, st>>>((const __nv_bfloat16*)dzd_bf16, y, y2, gamma, beta, stats, p,
                                                              seed, seed_salt, pg, pb, R, d);
    if (dgamma) colsum_final_kernel<<<ceil_div(d, 256), 256, 0, st>>>(pg, kColBlocks, d, dgamma, 0);
    if (dbeta) colsum_final_kernel<<<ceil_div(d, 256), 256, 0, st>>>(pb, kColBlocks, d, dbeta, 0);
  }
  B200_CUDA_LAUNCH_CHECK("layernorm_bwd");
  return B200_OK;
}

// dgamma / dbeta of the same backward, from the row statistics the call above left at the head of its workspace: a
// separate entry point so that a caller can keep it off the critical path (side stream) — only the optimizer needs it.
extern "C" int b200_layernorm_param_grads(const void* dzd_bf16, const float* y, const float* y2, const float* gamma,
                                          const float* beta, float p, unsigned long long seed,
                                          const unsigned long long* seed_salt, float* dgamma, float* dbeta, int R, int d,
                                          void* workspace, size_t workspace_bytes, b200_stream_t stream) {
  B200_CHECK_ARG(dzd_bf16 && y && y2 && gamma && beta && (dgamma || dbeta), "layernorm_param_grads: null tensor");
  B200_CHECK_ARG(R >= 0 && d > 0, "layernorm_param_grads: bad shape");
  B200_CHECK_ARG(workspace && workspace_bytes >= b200_layernorm_bwd_workspace_bytes(R, d), "layernorm_param_grads: workspace too small");
  if (R == 0) return B200_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const float* stats = (const float*)workspace;
  float* pg = (float*)((unsigned char*)workspace + align_up((size_t)R * 2 * 4, 256));
  float* pb = pg + (size_t)kColBlocks * d;
  dim3 grid(ceil_div(d, 256), kColBlocks);
  layernorm_param_grad_partial_kernel<<<grid, 256, 0, st>>>((const __nv_bfloat16*)dzd_bf16, y, y2, gamma, beta, stats, p, seed,
                                                            seed_salt, pg, pb, R, d);
  if (dgamma) colsum_final_kernel<<<ceil_div(d, 256), 256, 0, st>>>(pg, kColBlocks, d, dgamma, 0);
  if (dbeta) colsum_final_kernel<<<ceil_div(d, 256), 256, 0, st>>>(pb, kColBlocks, d, dbeta, 0);
  B200_CUDA_LAUNCH_CHECK("layernorm_param_grads");
  return B200_OK;
}

extern "C" int b200_text_attention_bwd(const void* dp1, const void* dp2, int ldp, const float* x, const float* attn,
                                       const float* vp, const float* dattn_ext, float* dx, int accumulate_dx, void* d_o,
                                       void* ds, int ldds, int R, int d, int L, b200_stream_t stream) {
  B200_CHECK_ARG(dp1 && dp2 && x && attn && vp && dx && d_o && ds, "text_attention_bwd: null tensor");
  B200_CHECK_ARG(R >= 0 && d > 0 && L > 0, "text_attention_bwd: bad shape");
  if (L > kAbMaxL || ldds > kAbMaxL || ldds < L || d % 8 != 0 || ldp % 8 != 0) {
    set_error("text_attention_bwd: unsupported shape (L=%d <= ldds=%d <= %d, d=%d %% 8, ldp=%d %% 8)", L, ldds, kAbMaxL, d, ldp);
    return B200_ERR_UNSUPPORTED;
  }
  if (R == 0) return B200_OK;
  text_attention_bwd_kernel<<<ceil_div(R, kAbRows), kAbThreads, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)dp1, (const __nv_bfloat16*)dp2, ldp, x, attn, vp, dattn_ext, dx, accumulate_dx,
      (__nv_bfloat16*)d_o, (__nv_bfloat16*)ds, ldds, R, d, L);
  B200_CUDA_LAUNCH_CHECK("text_attention_bwd");
  return B200_OK;
}

extern "C" int b200_head_losses(const float* logits, const float* deltas, const float* attn, const int64_t* gt_classes,
                                const float* proposals, const float* gt_boxes, int R, int K, int L, int cls_agnostic,
                                float wx, float wy, float ww, float wh, float smooth_l1_beta, float* out3, float* acc_stats5,
                                b200_stream_t stream) {
  B200_CHECK_ARG(logits && deltas && gt_classes && proposals && gt_boxes && out3, "head_losses: null tensor");
  B200_CHECK_ARG(R > 0 && K > 0 && (!attn || L > K), "head_losses: bad shape");
  head_losses_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(logits, deltas, attn, gt_classes, proposals, gt_boxes, R, K, L,
                                                           cls_agnostic, wx, wy, ww, wh, smooth_l1_beta, out3, acc_stats5);
  B200_CUDA_LAUNCH_CHECK("head_losses");
  return B200_OK;
}

extern "C" int b200_head_losses_bwd(const float* logits, const float* deltas, const float* attn, const int64_t* gt_classes,
                                    const float* proposals, const float* gt_boxes, const float* grad_scale3, int R, int K,
                                    int L, int cls_agnostic, float wx, float wy, float ww, float wh, float smooth_l1_beta,
                                    void* dlogits_bf16, int ldl, void* ddeltas_bf16, int ldd, float* dattn,
                                    b200_stream_t stream) {
  B200_CHECK_ARG(logits && deltas && gt_classes && proposals && gt_boxes && grad_scale3 && dlogits_bf16 && ddeltas_bf16,
                 "head_losses_bwd: null tensor");
  B200_CHECK_ARG(R > 0 && K > 0 && ldl >= K + 1 && ldd >= (cls_agnostic ? 4 : 4 * K), "head_losses_bwd: bad shape");
  head_losses_bwd_kernel<<<ceil_div(R, 256), 256, 0, (cudaStream_t)stream>>>(
      logits, deltas, attn, gt_classes, proposals, gt_boxes, grad_scale3, R, K, L, cls_agnostic, wx, wy, ww, wh,
      smooth_l1_beta, (__nv_bfloat16*)dlogits_bf16, ldl, (__nv_bfloat16*)ddeltas_bf16, ldd, dattn);
  B200_CUDA_LAUNCH_CHECK("head_losses_bwd");
  return B200_OK;
}

extern "C" int b200_sgd_momentum(float* params, const float* grads, float* momentum_buf, size_t n, float lr, float momentum,
                                 float weight_decay, void* bf16_shadow, b200_stream_t stream) {
  if (n == 0) return B200_OK;
  B200_CHECK_ARG(params && grads && momentum_buf, "sgd_momentum: null tensor");
  const int blocks = (int)min((size_t)kNumSMs * 8, (n + 255) / 256);
  sgd_momentum_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(params, grads, momentum_buf, n, lr, momentum, weight_decay,
                                                                (__nv_bfloat16*)bf16_shadow);
  B200_CUDA_LAUNCH_CHECK("sgd_momentum");
  return B200_OK;
}

extern "C" int b200_kd_loss(const float* student_logits, const float* teacher_logits, const int64_t* gt_classes, int R, int C1,
                            int bg_label, float temperature, float alpha, float* out1, b200_stream_t stream) {
  B200_CHECK_ARG(student_logits && teacher_logits && gt_classes && out1, "kd_loss: null tensor");
  B200_CHECK_ARG(R > 0 && C1 > 0 && temperature > 0.f, "kd_loss: bad shape / temperature");
  kd_loss_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(student_logits, teacher_logits, gt_classes, R, C1, bg_label, temperature,
                                                      alpha, out1);
  B200_CUDA_LAUNCH_CHECK("kd_loss");
  return B200_OK;
}

extern "C" int b200_kd_loss_bwd(const float* student_logits, const float* teacher_logits, const int64_t* gt_classes,
                                const float* grad_scale1, int R, int C1, int bg_label, float temperature, float alpha,
                                void* dlogits_bf16, int ldl, b200_stream_t stream) {
  B200_CHECK_ARG(student_logits && teacher_logits && gt_classes && grad_scale1 && dlogits_bf16, "kd_loss_bwd: null tensor");
  B200_CHECK_ARG(R > 0 && C1 > 0 && ldl >= C1 && temperature > 0.f, "kd_loss_bwd: bad shape / temperature");
  kd_loss_bwd_kernel<<<ceil_div(R, 256), 256, 0, (cudaStream_t)stream>>>(student_logits, teacher_logits, gt_classes, grad_scale1,
                                                                        R, C1, bg_label, temperature, alpha,
                                                                        (__nv_bfloat16*)dlogits_bf16, ldl);
  B200_CUDA_LAUNCH_CHECK("kd_loss_bwd");
  return B200_OK;
}
