// Non-GEMM pieces of the text-fusion chain (A3 core, A4 gate operands, A5 LayerNorm tail, A6 ReLU) and Q2 (PCB).
// Reference: defrcn/modeling/roi_heads/attentive_modules.py:45-55 (ScaledDotProductAttention),
// :166-174 (gating inputs), :71-75 (FFN residual + LayerNorm), :285 (ReLU);
// defrcn/evaluation/calibration_layer.py:110-123 (PCB cosine blend).
#include "common.cuh"

namespace b200 {

__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

// ------------------------------------------------------------------------------------------------
// text attention: the key set is tiny (L = K+2 <= 128 text keys), the query set is the ROIs.
// CTA = TR query rows.  Phase 1: S = q.Kp^T/sqrt(d) (warp per row-pair, lanes over d, Kp streamed from
// L2 once per CTA and shared by the warps through L1).  Phase 2: softmax per row.  Phase 3: O = attn.Vp
// with each thread owning 8 output columns x TR rows in registers; epilogue writes the two gate operands
// P1 = O*x and P2 = x-O as bf16 straight into the GEMM A-operand buffers (no fp32 round trip of O).
// ------------------------------------------------------------------------------------------------
constexpr int kAttThreads = 256;
constexpr int kAttMaxL = 128;

template <int TR>
__global__ void __launch_bounds__(kAttThreads)
text_attention_kernel(const __nv_bfloat16* __restrict__ q, const float* __restrict__ scores_in,
                      const void* __restrict__ xv, int x_is_bf16,
                      const float* __restrict__ kp, const float* __restrict__ vp, float* __restrict__ attn_out,
                      __nv_bfloat16* __restrict__ p1, __nv_bfloat16* __restrict__ p2, int ldp, int R, int d, int L,
                      float inv_temp) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  float* s_attn = reinterpret_cast<float*>(s_raw);           // [TR][kAttMaxL]
  float* s_q = s_attn + (size_t)TR * kAttMaxL;               // [TR][d] fp32 (only when q is given)
  const int r0 = blockIdx.x * TR;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwarps = kAttThreads / 32;

  if (scores_in) {
    // scores precomputed by the folded GEMM S = x (Kp Wq)^T / sqrt(d): skip phase 1
    for (int i = threadIdx.x; i < TR * L; i += kAttThreads) {
      const int row = i / L, l = i - row * L;
      s_attn[row * kAttMaxL + l] = (r0 + row < R) ? scores_in[(size_t)(r0 + row) * L + l] : 0.f;
    }
  } else {
  // stage q rows as fp32
  for (int i = threadIdx.x; i < TR * d / 2; i += kAttThreads) {
    const int row = (2 * i) / d, col = (2 * i) % d;
    const int r = r0 + row;
    uint32_t w = 0u;
    if (r < R) w = *reinterpret_cast<const uint32_t*>(q + (size_t)r * d + col);
    s_q[row * d + col] = bf16_lo(w);
    s_q[row * d + col + 1] = bf16_hi(w);
  }
  __syncthreads();

  // phase 1: scores.  all warps walk the keys in the same order so a key row is fetched from L2 once.
  constexpr int RPW = (TR + 7) / 8;  // rows per warp
  for (int l = 0; l < L; ++l) {
    const float4* krow = reinterpret_cast<const float4*>(kp + (size_t)l * d);
    float acc[RPW];
#pragma unroll
    for (int k = 0; k < RPW; ++k) acc[k] = 0.f;
    for (int i = lane; i < d / 4; i += 32) {
      const float4 kv = __ldg(krow + i);
#pragma unroll
      for (int k = 0; k < RPW; ++k) {
        const int row = warp + k * nwarps;
        if (row < TR) {
          const float4 qv = *reinterpret_cast<const float4*>(s_q + row * d + 4 * i);
          acc[k] += qv.x * kv.x + qv.y * kv.y + qv.z * kv.z + qv.w * kv.w;
        }
      }
    }
#pragma unroll
    for (int k = 0; k < RPW; ++k) {
      const float s = warp_sum(acc[k]);
      const int row = warp + k * nwarps;
      if (lane == 0 && row < TR) s_attn[row * kAttMaxL + l] = s * inv_temp;
    }
  }
  }
  __syncthreads();
  // phase 2: softmax over L (one warp per row)
  for (int row = warp; row < TR; row += nwarps) {
    float v[kAttMaxL / 32];
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < kAttMaxL / 32; ++j) {
      const int l = lane + 32 * j;
      v[j] = l < L ? s_attn[row * kAttMaxL + l] : -INFINITY;
      mx = fmaxf(mx, v[j]);
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < kAttMaxL / 32; ++j) {
      const int l = lane + 32 * j;
      v[j] = l < L ? expf(v[j] - mx) : 0.f;
      sum += v[j];
    }
    sum = warp_sum(sum);
    const int r = r0 + row;
#pragma unroll
    for (int j = 0; j < kAttMaxL / 32; ++j) {
      const int l = lane + 32 * j;
      if (l < L) {
        const float a = v[j] / sum;
        s_attn[row * kAttMaxL + l] = a;
        if (r < R && attn_out) attn_out[(size_t)r * L + l] = a;
      }
    }
  }
  __syncthreads();
  // phase 3: O = attn . Vp ; thread owns columns [c, c+8) per pass of kAttThreads*8 columns
  for (int c = threadIdx.x * 8; c < d; c += kAttThreads * 8) {
    float o[TR][8];
#pragma unroll
    for (int row = 0; row < TR; ++row)
#pragma unroll
      for (int k = 0; k < 8; ++k) o[row][k] = 0.f;
#pragma unroll 4
    for (int l = 0; l < L; ++l) {
      const float4 va = __ldg(reinterpret_cast<const float4*>(vp + (size_t)l * d + c));
      const float4 vb = __ldg(reinterpret_cast<const float4*>(vp + (size_t)l * d + c + 4));
      const float vv[8] = {va.x, va.y, va.z, va.w, vb.x, vb.y, vb.z, vb.w};
#pragma unroll
      for (int row = 0; row < TR; ++row) {
        const float a = s_attn[row * kAttMaxL + l];
#pragma unroll
        for (int k = 0; k < 8; ++k) o[row][k] += a * vv[k];
      }
    }
#pragma unroll
    for (int row = 0; row < TR; ++row) {
      const int r = r0 + row;
      if (r >= R) continue;
      float xx[8];
      if (x_is_bf16) {
        const uint4 t = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(xv) + (size_t)r * d + c);
        xx[0] = bf16_lo(t.x); xx[1] = bf16_hi(t.x); xx[2] = bf16_lo(t.y); xx[3] = bf16_hi(t.y);
        xx[4] = bf16_lo(t.z); xx[5] = bf16_hi(t.z); xx[6] = bf16_lo(t.w); xx[7] = bf16_hi(t.w);
      } else {
        const float4 a = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(xv) + (size_t)r * d + c);
        const float4 b = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(xv) + (size_t)r * d + c + 4);
        xx[0] = a.x; xx[1] = a.y; xx[2] = a.z; xx[3] = a.w; xx[4] = b.x; xx[5] = b.y; xx[6] = b.z; xx[7] = b.w;
      }
      uint4 w1, w2;
      w1.x = pack_bf16(o[row][0] * xx[0], o[row][1] * xx[1]); w1.y = pack_bf16(o[row][2] * xx[2], o[row][3] * xx[3]);
      w1.z = pack_bf16(o[row][4] * xx[4], o[row][5] * xx[5]); w1.w = pack_bf16(o[row][6] * xx[6], o[row][7] * xx[7]);
      w2.x = pack_bf16(xx[0] - o[row][0], xx[1] - o[row][1]); w2.y = pack_bf16(xx[2] - o[row][2], xx[3] - o[row][3]);
      w2.z = pack_bf16(xx[4] - o[row][4], xx[5] - o[row][5]); w2.w = pack_bf16(xx[6] - o[row][6], xx[7] - o[row][7]);
      *reinterpret_cast<uint4*>(p1 + (size_t)r * ldp + c) = w1;
      *reinterpret_cast<uint4*>(p2 + (size_t)r * ldp + c) = w2;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// residual + LayerNorm (+ReLU): one warp per row, two-pass statistics in fp32 (mean, then centred variance)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
residual_layernorm_kernel(const float* __restrict__ y, const float* __restrict__ y2, const float* __restrict__ gamma,
                          const float* __restrict__ beta, float eps, int relu, float* __restrict__ out_f32,
                          __nv_bfloat16* __restrict__ out_bf16, int R, int d) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= R) return;
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
    float4 o;
    o.x = (v.x - mean) * rstd * g.x + be.x; o.y = (v.y - mean) * rstd * g.y + be.y;
    o.z = (v.z - mean) * rstd * g.z + be.z; o.w = (v.w - mean) * rstd * g.w + be.w;
    if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
    if (out_f32) reinterpret_cast<float4*>(out_f32 + (size_t)row * d)[i] = o;
    if (out_bf16) {
      uint2 w;
      w.x = pack_bf16(o.x, o.y); w.y = pack_bf16(o.z, o.w);
      reinterpret_cast<uint2*>(out_bf16 + (size_t)row * d)[i] = w;
    }
  }
}

__global__ void cast_bf16_kernel(const float* __restrict__ src, int ld_src, __nv_bfloat16* __restrict__ dst, int ld_dst,
                                 int rows, int cols4) {
  const size_t total = (size_t)rows * cols4;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / cols4), c = (int)(i % cols4) * 4;
    const float4 v = *reinterpret_cast<const float4*>(src + (size_t)r * ld_src + c);
    uint2 w;
    w.x = pack_bf16(v.x, v.y); w.y = pack_bf16(v.z, v.w);
    *reinterpret_cast<uint2*>(dst + (size_t)r * ld_dst + c) = w;
  }
}

// ------------------------------------------------------------------------------------------------
// PCB: a single CTA (n <= ~100 detections per image): the [ileft, iright) window is counted from the
// ORIGINAL scores before any of them is rewritten, then one warp per detection.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
pcb_cosine_blend_kernel(float* __restrict__ scores, const float* __restrict__ feats, const float* __restrict__ protos,
                        const int64_t* __restrict__ classes, const uint8_t* __restrict__ exclude, int n, int D, int K,
                        float alpha, float lower, float upper) {
  __shared__ int s_left, s_right;
  if (threadIdx.x == 0) { s_left = 0; s_right = 0; }
  __syncthreads();
  int l = 0, r = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float s = scores[i];
    l += s > upper;
    r += s > lower;
  }
  l = (int)warp_sum((float)l); r = (int)warp_sum((float)r);
  if ((threadIdx.x & 31) == 0) { atomicAdd(&s_left, l); atomicAdd(&s_right, r); }
  __syncthreads();
  const int ileft = s_left, iright = s_right;
  __syncthreads();  // everyone has read the ORIGINAL scores' counts before any score is modified by this block
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = ileft + warp; i < iright; i += (blockDim.x >> 5)) {
    const int c = (int)classes[i];
    if (c < 0 || c >= K || (exclude && exclude[c])) continue;
    const float* f = feats + (size_t)i * D;
    const float* p = protos + (size_t)c * D;
    float dot = 0.f, nf = 0.f, np = 0.f;
    for (int k = lane; k < D; k += 32) {
      const float a = f[k], b = p[k];
      dot += a * b; nf += a * a; np += b * b;
    }
    dot = warp_sum(dot); nf = sqrtf(warp_sum(nf)); np = sqrtf(warp_sum(np));
    if (nf == 0.f) nf = 1.f;
    if (np == 0.f) np = 1.f;
    const float cosv = dot / (nf * np);
    if (lane == 0) scores[i] = scores[i] * alpha + cosv * (1.f - alpha);
  }
}

}  // namespace b200

using namespace b200;

template <int TR>
static int launch_att(const void* q, const float* scores_in, const void* x, int x_dtype, const float* kp, const float* vp,
                      float* attn_out, void* p1, void* p2, int ldp, int R, int d, int L, cudaStream_t st) {
  const size_t smem = (q ? (size_t)TR * d * 4 : 0) + (size_t)TR * kAttMaxL * 4;
  auto k = text_attention_kernel<TR>;
  B200_CUDA_CALL(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k<<<ceil_div(R, TR), kAttThreads, smem, st>>>((const __nv_bfloat16*)q, scores_in, x, x_dtype == B200_BF16, kp, vp, attn_out,
                                                (__nv_bfloat16*)p1, (__nv_bfloat16*)p2, ldp, R, d, L,
                                                1.0f / sqrtf((float)d));
  B200_CUDA_LAUNCH_CHECK("text_attention");
  return B200_OK;
}

extern "C" int b200_text_attention(const void* q, const float* scores_in, const void* x, int x_dtype, const float* kp,
                                   const float* vp, float* attn_out, void* p1, void* p2, int ldp, int R, int d, int L,
                                   b200_stream_t stream) {
  B200_CHECK_ARG((q || scores_in) && x && vp && p1 && p2 && (kp || !q), "text_attention: null tensor");
  B200_CHECK_ARG(R >= 0 && d > 0 && L > 0, "text_attention: bad shape");
  if (L > kAttMaxL || d % 8 != 0 || ldp % 8 != 0 || (size_t)8 * d * 4 > 200 * 1024) {
    set_error("text_attention: unsupported shape (L=%d<=%d, d=%d %% 8, ldp=%d %% 8)", L, kAttMaxL, d, ldp);
    return B200_ERR_UNSUPPORTED;
  }
  if (R == 0) return B200_OK;
  cudaStream_t st = (cudaStream_t)stream;
  // enough CTAs to cover the SMs at small R, fewer Kp/Vp re-reads from L2 at large R
  if (R <= 148 * 4) return launch_att<4>(q, scores_in, x, x_dtype, kp, vp, attn_out, p1, p2, ldp, R, d, L, st);
  return launch_att<8>(q, scores_in, x, x_dtype, kp, vp, attn_out, p1, p2, ldp, R, d, L, st);
}

extern "C" int b200_residual_layernorm(const float* y, const float* y2, const float* gamma, const float* beta, float eps,
                                       int relu, float* out_f32, void* out_bf16, int R, int d, b200_stream_t stream) {
  B200_CHECK_ARG(y && gamma && beta && (out_f32 || out_bf16), "residual_layernorm: null tensor");
  B200_CHECK_ARG(R >= 0 && d > 0 && d % 4 == 0, "residual_layernorm: d must be a multiple of 4");
  if (R == 0) return B200_OK;
  residual_layernorm_kernel<<<ceil_div(R, 8), 256, 0, (cudaStream_t)stream>>>(y, y2, gamma, beta, eps, relu, out_f32,
                                                                              (__nv_bfloat16*)out_bf16, R, d);
  B200_CUDA_LAUNCH_CHECK("residual_layernorm");
  return B200_OK;
}

// Optional cosine + temperature form of the prototype logits (C1'): rows scaled to unit L2 norm (norm clamped at eps, the
// reference helper's rule: my_module.py:461-469 `sim_matrix`), times `scale`, as the bf16 operand of the logits GEMM.
// One warp per row, two passes over a row that stays in L1.
template <typename T>
__global__ void __launch_bounds__(256)
l2_normalize_rows_kernel(const T* __restrict__ src, int ld_src, __nv_bfloat16* __restrict__ dst, int ld_dst, int rows, int cols,
                         float eps, float scale) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const T* s = src + (size_t)row * ld_src;
  float ss = 0.f;
  for (int c = lane; c < cols; c += 32) {
    const float v = (float)s[c];
    ss += v * v;
  }
  ss = warp_sum(ss);
  const float inv = scale / fmaxf(sqrtf(ss), eps);
  __nv_bfloat16* d = dst + (size_t)row * ld_dst;
  for (int c = lane; c < cols; c += 32) d[c] = __float2bfloat16_rn((float)s[c] * inv);
}

extern "C" int b200_l2_normalize_rows(const void* src, int src_dtype, int ld_src, void* dst_bf16, int ld_dst, int rows, int cols,
                                      float eps, float scale, b200_stream_t stream) {
  B200_CHECK_ARG(src && dst_bf16, "l2_normalize_rows: null tensor");
  B200_CHECK_ARG(rows >= 0 && cols > 0 && ld_src >= cols && ld_dst >= cols && (src_dtype | 1) == 1, "l2_normalize_rows: bad shape");
  if (rows == 0) return B200_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (src_dtype == B200_F32)
    l2_normalize_rows_kernel<float><<<ceil_div(rows, 8), 256, 0, st>>>((const float*)src, ld_src, (__nv_bfloat16*)dst_bf16, ld_dst, rows, cols, eps, scale);
  else
    l2_normalize_rows_kernel<__nv_bfloat16><<<ceil_div(rows, 8), 256, 0, st>>>((const __nv_bfloat16*)src, ld_src, (__nv_bfloat16*)dst_bf16, ld_dst, rows, cols, eps, scale);
  B200_CUDA_LAUNCH_CHECK("l2_normalize_rows");
  return B200_OK;
}

// A7 (teacher attention): dst[r][:] = bf16(act(table[idx[r]][:])) — the per-ROI text feature of the teachers is a row of
// the (K+1)-row class table picked by the ground-truth label (attentive_modules.py:380-401: one_hot(label) @ table)
__global__ void __launch_bounds__(256)
gather_rows_bf16_kernel(const float* __restrict__ table, int ld_table, const int64_t* __restrict__ idx, int num_rows_table,
                        __nv_bfloat16* __restrict__ dst, int ld_dst, int rows, int cols4, int relu) {
  const size_t total = (size_t)rows * cols4;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / cols4), c = (int)(i % cols4) * 4;
    long long t = idx[r];
    t = t < 0 ? 0 : (t >= num_rows_table ? num_rows_table - 1 : t);
    float4 v = *reinterpret_cast<const float4*>(table + (size_t)t * ld_table + c);
    if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
    __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
    uint2 o;
    o.x = *reinterpret_cast<uint32_t*>(&lo);
    o.y = *reinterpret_cast<uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(dst + (size_t)r * ld_dst + c) = o;
  }
}

// A7 (teacher attention): per-class mean of feature rows and class counts.  The teachers attend over one key / value per
// ROI of the batch, but ROIs of one class share their key, so softmax(Q K^T) V over R + 1 keys equals a (K+2)-key attention
// whose class-c logit carries + log n_c and whose class-c value is the MEAN of the class's value rows (see
// ops.teacher_attention_forward).  Two fixed-order stages: 128-row chunks accumulate per-class partial sums in shared
// memory (a thread owns a column: no conflicts, no atomics), then the chunks are summed in order and divided by n_c.
constexpr int kCmRows = 128, kCmCols = 256;
__global__ void __launch_bounds__(kCmCols)
class_sum_partial_kernel(const float* __restrict__ x, int ldx, const int64_t* __restrict__ labels, int R, int d, int C,
                         float* __restrict__ partial /*[chunks][C][d]*/) {
  extern __shared__ float s_acc[];                 // [C][kCmCols]
  __shared__ int s_lab[kCmRows];
  const int chunk = blockIdx.y, col = blockIdx.x * kCmCols + threadIdx.x;
  const int r0 = chunk * kCmRows, nr = min(kCmRows, R - r0);
  for (int c = 0; c < C; ++c) s_acc[c * kCmCols + threadIdx.x] = 0.f;
  if ((int)threadIdx.x < nr) {
    long long l = labels[r0 + threadIdx.x];
    s_lab[threadIdx.x] = (int)(l < 0 ? 0 : (l >= C ? C - 1 : l));
  }
  __syncthreads();
  if (col < d)
    for (int r = 0; r < nr; ++r) s_acc[s_lab[r] * kCmCols + threadIdx.x] += x[(size_t)(r0 + r) * ldx + col];
  if (col < d)
    for (int c = 0; c < C; ++c) partial[((size_t)chunk * C + c) * d + col] = s_acc[c * kCmCols + threadIdx.x];
}

__global__ void __launch_bounds__(256)
class_mean_final_kernel(const float* __restrict__ partial, const int64_t* __restrict__ labels, int R, int d, int C, int chunks,
                        float* __restrict__ mean /*[C][d]*/, float* __restrict__ counts /*[C]*/) {
  __shared__ int s_cnt;
  const int c = blockIdx.y;
  if (threadIdx.x == 0) s_cnt = 0;
  __syncthreads();
  int local = 0;
  for (int r = threadIdx.x; r < R; r += blockDim.x) {
    long long l = labels[r];
    l = l < 0 ? 0 : (l >= C ? C - 1 : l);
    local += (l == c);
  }
  if (local) atomicAdd(&s_cnt, local);             // integer: order independent
  __syncthreads();
  const int n = s_cnt;
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col < d) {
    float acc = 0.f;
    for (int k = 0; k < chunks; ++k) acc += partial[((size_t)k * C + c) * d + col];
    mean[(size_t)c * d + col] = n > 0 ? acc / (float)n : 0.f;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) counts[c] = (float)n;
}

extern "C" size_t b200_class_mean_rows_workspace_bytes(int R, int d, int C) {
  if (R <= 0 || d <= 0 || C <= 0) return 256;
  return align_up((size_t)ceil_div(R, kCmRows) * C * d * sizeof(float), 256);
}

extern "C" int b200_class_mean_rows(const float* x, int ldx, const int64_t* labels, int R, int d, int C, float* mean,
                                    float* counts, void* workspace, size_t workspace_bytes, b200_stream_t stream) {
  B200_CHECK_ARG(x && labels && mean && counts, "class_mean_rows: null tensor");
  B200_CHECK_ARG(R > 0 && d > 0 && C > 0 && (size_t)C * kCmCols * sizeof(float) <= 200 * 1024, "class_mean_rows: bad shape");
  if (!workspace || workspace_bytes < b200_class_mean_rows_workspace_bytes(R, d, C)) {
    set_error("class_mean_rows: workspace too small");
    return B200_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int chunks = ceil_div(R, kCmRows);
  const size_t smem = (size_t)C * kCmCols * sizeof(float);
  if (smem > 48 * 1024)
    B200_CUDA_CALL(cudaFuncSetAttribute(class_sum_partial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  class_sum_partial_kernel<<<dim3(ceil_div(d, kCmCols), chunks), kCmCols, smem, st>>>(x, ldx, labels, R, d, C, (float*)workspace);
  B200_CUDA_LAUNCH_CHECK("class_mean_rows(partial)");
  class_mean_final_kernel<<<dim3(ceil_div(d, 256), C), 256, 0, st>>>((const float*)workspace, labels, R, d, C, chunks, mean, counts);
  B200_CUDA_LAUNCH_CHECK("class_mean_rows(final)");
  return B200_OK;
}

extern "C" int b200_gather_rows_bf16(const float* table, int ld_table, int num_rows_table, const int64_t* idx, void* dst,
                                     int ld_dst, int rows, int cols, int relu, b200_stream_t stream) {
  B200_CHECK_ARG(table && idx && dst, "gather_rows_bf16: null tensor");
  B200_CHECK_ARG(rows >= 0 && cols >= 0 && num_rows_table > 0 && cols % 4 == 0 && ld_table % 4 == 0 && ld_dst % 4 == 0,
                 "gather_rows_bf16: cols/ld must be multiples of 4");
  if (rows == 0 || cols == 0) return B200_OK;
  const size_t total = (size_t)rows * (cols / 4);
  const int blocks = (int)min((size_t)kNumSMs * 8, (total + 255) / 256);
  gather_rows_bf16_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(table, ld_table, idx, num_rows_table,
                                                                   (__nv_bfloat16*)dst, ld_dst, rows, cols / 4, relu);
  B200_CUDA_LAUNCH_CHECK("gather_rows_bf16");
  return B200_OK;
}

extern "C" int b200_cast_bf16(const float* src, int ld_src, void* dst, int ld_dst, int rows, int cols,
                              b200_stream_t stream) {
  B200_CHECK_ARG(src && dst, "cast_bf16: null tensor");
  B200_CHECK_ARG(rows >= 0 && cols >= 0 && cols % 4 == 0 && ld_src % 4 == 0 && ld_dst % 4 == 0,
                 "cast_bf16: cols/ld must be multiples of 4");
  if (rows == 0 || cols == 0) return B200_OK;
  const size_t total = (size_t)rows * (cols / 4);
  const int blocks = (int)min((size_t)kNumSMs * 8, (total + 255) / 256);
  cast_bf16_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(src, ld_src, (__nv_bfloat16*)dst, ld_dst, rows, cols / 4);
  B200_CUDA_LAUNCH_CHECK("cast_bf16");
  return B200_OK;
}

extern "C" int b200_pcb_cosine_blend(float* scores, const float* feats, const float* prototypes, const int64_t* classes,
                                     const uint8_t* exclude, int n, int D, int K, float alpha, float lower, float upper,
                                     b200_stream_t stream) {
  B200_CHECK_ARG(n >= 0 && D > 0 && K > 0, "pcb_cosine_blend: bad shape");
  if (n == 0) return B200_OK;
  B200_CHECK_ARG(scores && feats && prototypes && classes, "pcb_cosine_blend: null tensor");
  pcb_cosine_blend_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(scores, feats, prototypes, classes, exclude,
                                                                            n, D, K, alpha, lower, upper);
  B200_CUDA_LAUNCH_CHECK("pcb_cosine_blend");
  return B200_OK;
}
