// T1 + the text half of A1/A2: the (K+2)-row text side of the fusion attention, forward and backward, in fp32.
// Reference: defrcn/modeling/roi_heads/attentive_modules.py:274-277 (key/value projection + ReLU of the class-name
// embeddings), :125-135 (w_k / w_v, dummy key, zero value) and the folded query operand Kq = Kp Wq / sqrt(d).
//
// All of these are contractions with at most 32 rows on one side (the text matrix has K+1 <= 81 rows in general but
// the head's tables are built per <= 32-row block): tensor-core tiles would be > 80 % padding and cuBLAS falls back
// to SIMT sgemm kernels that take 30-50 us each for < 0.2 GFLOP (ncu launch list, round 1).  Three small fp32 kernels
// cover every product of the forward and of autograd's backward; each streams the big operand (the 2048 x 2048 or
// 2048 x D weight) once, coalesced, and is bound by that read:
//   NT  out[m][n] = act(sum_k A[m][k] B[n][k] + bias[n])          linear forward          (B = weight)
//   NN  out[m][k] = scale * sum_n A'[m][n] B[n][k]                 linear data gradient    (B = weight), Kq = Kp Wq
//   TN  out[n][k] = sum_m A'[m][n] B[m][k],  ob[n] = sum_m A'[m][n]  linear weight / bias gradient
// A' = A masked by (ref > 0) when a ReLU sits between (ref = the forward activation).  Deterministic: NN splits the
// n-range over CTAs into partials that are summed in a fixed order.
#include "common.cuh"

namespace b200 {

constexpr int kSkMaxM = 32;

__device__ __forceinline__ float masked(const float* a, const float* ref, size_t i) {
  const float v = a[i];
  return (ref && !(ref[i] > 0.f)) ? 0.f : v;
}

// ---- NT: warp <-> 2 output columns, lanes over k ------------------------------------------------------------------
template <int MM>
__global__ void __launch_bounds__(256)
skinny_nt_kernel(const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb, const float* __restrict__ bias,
                 int relu, float* __restrict__ out, int ldo, int M, int N, int K) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = (blockIdx.x * 8 + warp) * 2;
  if (n0 >= N) return;
  const bool two = n0 + 1 < N;
  float acc0[MM], acc1[MM];
#pragma unroll
  for (int m = 0; m < MM; ++m) { acc0[m] = 0.f; acc1[m] = 0.f; }
  const float* b0 = B + (size_t)n0 * ldb;
  const float* b1 = B + (size_t)(two ? n0 + 1 : n0) * ldb;
  for (int k = lane * 4; k < K; k += 128) {
    const float4 w0 = __ldg(reinterpret_cast<const float4*>(b0 + k));
    const float4 w1 = __ldg(reinterpret_cast<const float4*>(b1 + k));
#pragma unroll
    for (int m = 0; m < MM; ++m) {
      if (m < M) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(A + (size_t)m * lda + k));
        acc0[m] += a.x * w0.x + a.y * w0.y + a.z * w0.z + a.w * w0.w;
        acc1[m] += a.x * w1.x + a.y * w1.y + a.z * w1.z + a.w * w1.w;
      }
    }
  }
#pragma unroll
  for (int m = 0; m < MM; ++m) {
    if (m < M) {
      float s0 = warp_sum(acc0[m]), s1 = warp_sum(acc1[m]);
      if (lane == 0) {
        s0 += bias ? bias[n0] : 0.f;
        out[(size_t)m * ldo + n0] = relu ? fmaxf(s0, 0.f) : s0;
        if (two) {
          s1 += bias ? bias[n0 + 1] : 0.f;
          out[(size_t)m * ldo + n0 + 1] = relu ? fmaxf(s1, 0.f) : s1;
        }
      }
    }
  }
}

// ---- NN: thread <-> output column k, CTAs split the n-range; partial[split][m][k] then an ordered sum --------------
constexpr int kNnSplit = 16;

template <int MM>
__global__ void __launch_bounds__(256)
skinny_nn_partial_kernel(const float* __restrict__ A, int lda, const float* __restrict__ ref, int ldref,
                         const float* __restrict__ B, int ldb, float* __restrict__ partial, int M, int N, int K) {
  __shared__ float s_a[MM][64];
  const int k = blockIdx.x * 256 + threadIdx.x;
  const int per = (N + kNnSplit - 1) / kNnSplit;
  const int nb = blockIdx.y * per, ne = min(N, nb + per);
  float acc[MM];
#pragma unroll
  for (int m = 0; m < MM; ++m) acc[m] = 0.f;
  for (int n0 = nb; n0 < ne; n0 += 64) {
    __syncthreads();
    for (int i = threadIdx.x; i < MM * 64; i += 256) {
      const int m = i >> 6, j = i & 63;
      float v = 0.f;
      if (m < M && n0 + j < ne) {
        v = A[(size_t)m * lda + n0 + j];
        if (ref && !(ref[(size_t)m * ldref + n0 + j] > 0.f)) v = 0.f;
      }
      s_a[m][j] = v;
    }
    __syncthreads();
    if (k < K) {
      const int lim = min(64, ne - n0);
      for (int j = 0; j < lim; ++j) {
        const float w = __ldg(B + (size_t)(n0 + j) * ldb + k);
#pragma unroll
        for (int m = 0; m < MM; ++m) acc[m] += s_a[m][j] * w;
      }
    }
  }
  if (k < K) {
#pragma unroll
    for (int m = 0; m < MM; ++m)
      if (m < M) partial[((size_t)blockIdx.y * M + m) * K + k] = acc[m];
  }
}

__global__ void skinny_nn_final_kernel(const float* __restrict__ partial, float scale, float* __restrict__ out, int ldo, int M,
                                       int K) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * K) return;
  const int m = i / K, k = i - m * K;
  float s = 0.f;
  for (int sp = 0; sp < kNnSplit; ++sp) s += partial[((size_t)sp * M + m) * K + k];
  out[(size_t)m * ldo + k] = s * scale;
}

// ---- TN: thread <-> (n, 4 consecutive k) --------------------------------------------------------------------------
template <int MM>
__global__ void __launch_bounds__(256)
skinny_tn_kernel(const float* __restrict__ A, int lda, const float* __restrict__ ref, int ldref, const float* __restrict__ B,
                 int ldb, float* __restrict__ out, int ldo, float* __restrict__ out_bias, int M, int N, int K, int accumulate) {
  const int k = (blockIdx.x * 64 + (threadIdx.x & 63)) * 4;
  const int n = blockIdx.y * 4 + (threadIdx.x >> 6);
  if (n >= N || k >= K) return;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  float bsum = 0.f;
#pragma unroll
  for (int m = 0; m < MM; ++m) {
    if (m < M) {
      float a = A[(size_t)m * lda + n];
      if (ref && !(ref[(size_t)m * ldref + n] > 0.f)) a = 0.f;
      const float4 b = __ldg(reinterpret_cast<const float4*>(B + (size_t)m * ldb + k));
      acc.x += a * b.x; acc.y += a * b.y; acc.z += a * b.z; acc.w += a * b.w;
      bsum += a;
    }
  }
  float4* dst = reinterpret_cast<float4*>(out + (size_t)n * ldo + k);
  if (accumulate) {                       // row blocks of a taller A are summed in call order
    const float4 p = *dst;
    acc.x += p.x; acc.y += p.y; acc.z += p.z; acc.w += p.w;
  }
  *dst = acc;
  if (out_bias && k == 0) out_bias[n] = accumulate ? out_bias[n] + bsum : bsum;
}

template <int MM>
static int launch_skinny(int mode, const float* A, int lda, const float* ref, int ldref, const float* B, int ldb,
                         const float* bias, int relu, float scale, float* out, int ldo, float* out_bias, int M, int N, int K,
                         int accumulate, float* ws, cudaStream_t st) {
  if (mode == 0) {
    skinny_nt_kernel<MM><<<ceil_div(N, 16), 256, 0, st>>>(A, lda, B, ldb, bias, relu, out, ldo, M, N, K);
  } else if (mode == 1) {
    dim3 grid(ceil_div(K, 256), kNnSplit);
    skinny_nn_partial_kernel<MM><<<grid, 256, 0, st>>>(A, lda, ref, ldref, B, ldb, ws, M, N, K);
    skinny_nn_final_kernel<<<ceil_div(M * K, 256), 256, 0, st>>>(ws, scale, out, ldo, M, K);
  } else {
    dim3 grid(ceil_div(K, 256), ceil_div(N, 4));
    skinny_tn_kernel<MM><<<grid, 256, 0, st>>>(A, lda, ref, ldref, B, ldb, out, ldo, out_bias, M, N, K, accumulate);
  }
  B200_CUDA_LAUNCH_CHECK("skinny_gemm");
  return B200_OK;
}

}  // namespace b200

using namespace b200;

extern "C" size_t b200_skinny_gemm_workspace_bytes(int M, int K) { return (size_t)kNnSplit * max(M, 1) * max(K, 1) * 4; }

extern "C" int b200_skinny_gemm(int mode, const float* A, int lda, const float* relu_ref, int ldref, const float* B, int ldb,
                                const float* bias, int relu, float scale, float* out, int ldo, float* out_bias, int M,
                                int N, int K, int accumulate, void* workspace, size_t workspace_bytes,
                                b200_stream_t stream) {
  B200_CHECK_ARG(mode >= 0 && mode <= 2 && A && B && out, "skinny_gemm: bad mode or null tensor");
  B200_CHECK_ARG(M > 0 && M <= kSkMaxM && N > 0 && K > 0, "skinny_gemm: need 0 < M <= 32");
  B200_CHECK_ARG(!accumulate || mode == 2, "skinny_gemm: accumulate is a TN-mode option");
  if (K % 4 || ldb % 4 || ((uintptr_t)B & 15) || (mode == 0 && (lda % 4 || ((uintptr_t)A & 15))) ||
      (mode == 2 && (ldo % 4 || ((uintptr_t)out & 15)))) {
    set_error("skinny_gemm: K and the leading dimensions of the vector-accessed operands must be multiples of 4 floats");
    return B200_ERR_UNSUPPORTED;
  }
  if (mode == 1) B200_CHECK_ARG(workspace && workspace_bytes >= b200_skinny_gemm_workspace_bytes(M, K), "skinny_gemm: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  float* ws = (float*)workspace;
  if (M <= 8) return launch_skinny<8>(mode, A, lda, relu_ref, ldref, B, ldb, bias, relu, scale, out, ldo, out_bias, M, N, K, accumulate, ws, st);
  if (M <= 24) return launch_skinny<24>(mode, A, lda, relu_ref, ldref, B, ldb, bias, relu, scale, out, ldo, out_bias, M, N, K, accumulate, ws, st);
  return launch_skinny<32>(mode, A, lda, relu_ref, ldref, B, ldb, bias, relu, scale, out, ldo, out_bias, M, N, K, accumulate, ws, st);
}
