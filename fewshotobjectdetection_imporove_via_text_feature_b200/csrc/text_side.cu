// T1 + the text half of A1/A2: the (K+2)-row text side of the fusion attention, forward and backward, in fp32.
// Reference: defrcn/modeling/roi_heads/attentive_modules.py:274-277 (key/value projection + ReLU of the class-name
// embeddings), :125-135 (w_k / w_v, dummy key, zero value) and the folded query operand Kq = Kp Wq / sqrt(d).
//
// All of these are contractions with at most 32 rows on one side (the text matrix has K+1 <= 81 rows in general; the
// host feeds taller ones in 32-row blocks): tensor-core tiles would be > 80 % padding and cuBLAS falls back to SIMT
// sgemm kernels that take 30-50 us each for < 0.2 GFLOP (ncu launch list, round 1).  What bounds them is streaming the
// big operand (a 2048 x 2048 or 2048 x D fp32 weight, 16.8 MB) once, so the kernels are organised for parallelism
// and bytes in flight, not FLOPs:
//   NT  out[m][n] = act(sum_k A[m][k] B[n][k] + bias[n])            linear forward          (B = weight)
//   NN  out[m][k] = scale * sum_n A'[m][n] B[n][k]                   linear data gradient    (B = weight), Kq = Kp Wq
//   TN  out[n][k] = sum_m A'[m][n] B[m][k],  ob[n] = sum_m A'[m][n]   linear weight / bias gradient
// A' = A masked by (ref > 0) when a ReLU sits between (ref = the forward activation).
// NT / NN: a 32 x 64 output tile per CTA with the reduction split over up to 16 CTAs (512 CTAs for a 2048 x 2048
// weight, each streaming a 32 KB slab through padded shared-memory tiles, 2 x 4 register blocking); partials are
// summed in a fixed order by a small second kernel — deterministic.  TN: 4 x 4 outputs per thread, 22 FMAs each,
// bound by the 16.8 MB write.
#include "common.cuh"

namespace b200 {

constexpr int kSkMaxM = 32;
constexpr int kSkCols = 64;        // output columns per CTA (NT / NN)
constexpr int kSkRc = 64;          // reduction chunk staged per iteration
constexpr int kSkPad = 4;          // row padding (floats) of the shared tiles: conflict-free float4 reads across rows
constexpr int kSkMaxSplit = 16;

__host__ __device__ inline int sk_split(int reduce) { return max(1, min(kSkMaxSplit, reduce / 128)); }

// MODE 0 (NT): cols = n, reduce = k, B[col][r].  MODE 1 (NN): cols = k, reduce = n, B[r][col].
// grid (ceil(cols / 64), split), 256 threads: thread (tm, tc) owns rows {2tm, 2tm+1} x 4 columns.
template <int MODE>
__global__ void __launch_bounds__(256)
skinny_reduce_kernel(const float* __restrict__ A, int lda, const float* __restrict__ ref, int ldref,
                     const float* __restrict__ B, int ldb, float* __restrict__ partial, int M, int cols, int reduce) {
  __shared__ __align__(16) float s_a[kSkMaxM][kSkRc + kSkPad];
  __shared__ __align__(16) float s_b[kSkCols][kSkRc + kSkPad];      // MODE 0: [col][r]   MODE 1: [r][col] (kSkRc == kSkCols)
  const int tid = threadIdx.x, tm = tid >> 4, tc = tid & 15;
  const int c0 = blockIdx.x * kSkCols;
  const int S = gridDim.y;
  const int per = ((reduce + S - 1) / S + 3) & ~3;                  // multiple of 4: float4 loads stay aligned
  const int rb = blockIdx.y * per, re = min(reduce, rb + per);
  float acc[2][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int r0 = rb; r0 < re; r0 += kSkRc) {
    __syncthreads();
    // A tile: rows m, kSkRc reduction entries (zero beyond M / re), ReLU mask applied
    for (int i = tid; i < kSkMaxM * (kSkRc / 4); i += 256) {
      const int m = i / (kSkRc / 4), q = (i - m * (kSkRc / 4)) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m < M && r0 + q < re) {
        v = __ldg(reinterpret_cast<const float4*>(A + (size_t)m * lda + r0 + q));
        if (ref) {
          const float4 f = __ldg(reinterpret_cast<const float4*>(ref + (size_t)m * ldref + r0 + q));
          if (!(f.x > 0.f)) v.x = 0.f;
          if (!(f.y > 0.f)) v.y = 0.f;
          if (!(f.z > 0.f)) v.z = 0.f;
          if (!(f.w > 0.f)) v.w = 0.f;
        }
        if (r0 + q + 1 >= re) v.y = 0.f;
        if (r0 + q + 2 >= re) v.z = 0.f;
        if (r0 + q + 3 >= re) v.w = 0.f;
      }
      *reinterpret_cast<float4*>(&s_a[m][q]) = v;
    }
    // B tile, 64 x 64, one coalesced 256-byte row segment per 16 threads
    for (int i = tid; i < kSkCols * (kSkRc / 4); i += 256) {
      const int row = i / (kSkRc / 4), q = (i - row * (kSkRc / 4)) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (MODE == 0) {          // row = output column n, q = reduction offset
        if (c0 + row < cols && r0 + q < re) {
          v = __ldg(reinterpret_cast<const float4*>(B + (size_t)(c0 + row) * ldb + r0 + q));
          if (r0 + q + 1 >= re) v.y = 0.f;
          if (r0 + q + 2 >= re) v.z = 0.f;
          if (r0 + q + 3 >= re) v.w = 0.f;
        }
      } else {                  // row = reduction offset, q = output column offset
        if (r0 + row < re && c0 + q < cols) {
          v = __ldg(reinterpret_cast<const float4*>(B + (size_t)(r0 + row) * ldb + c0 + q));
          if (c0 + q + 1 >= cols) v.y = 0.f;
          if (c0 + q + 2 >= cols) v.z = 0.f;
          if (c0 + q + 3 >= cols) v.w = 0.f;
        }
      }
      *reinterpret_cast<float4*>(&s_b[row][q]) = v;
    }
    __syncthreads();
    if (MODE == 0) {            // columns tc + 16 j: rows of s_b 16 apart land on distinct banks
#pragma unroll 4
      for (int q = 0; q < kSkRc; q += 4) {
        const float4 a0 = *reinterpret_cast<const float4*>(&s_a[2 * tm][q]);
        const float4 a1 = *reinterpret_cast<const float4*>(&s_a[2 * tm + 1][q]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 b = *reinterpret_cast<const float4*>(&s_b[tc + 16 * j][q]);
          acc[0][j] += a0.x * b.x + a0.y * b.y + a0.z * b.z + a0.w * b.w;
          acc[1][j] += a1.x * b.x + a1.y * b.y + a1.z * b.z + a1.w * b.w;
        }
      }
    } else {                    // columns 4 tc .. 4 tc + 3
#pragma unroll 8
      for (int r = 0; r < kSkRc; ++r) {
        const float a0 = s_a[2 * tm][r], a1 = s_a[2 * tm + 1][r];
        const float4 b = *reinterpret_cast<const float4*>(&s_b[r][4 * tc]);
        acc[0][0] += a0 * b.x; acc[0][1] += a0 * b.y; acc[0][2] += a0 * b.z; acc[0][3] += a0 * b.w;
        acc[1][0] += a1 * b.x; acc[1][1] += a1 * b.y; acc[1][2] += a1 * b.z; acc[1][3] += a1 * b.w;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int m = 2 * tm + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = c0 + (MODE == 0 ? tc + 16 * j : 4 * tc + j);
      if (c < cols) partial[((size_t)blockIdx.y * M + m) * cols + c] = acc[i][j];
    }
  }
}

__global__ void skinny_final_kernel(const float* __restrict__ partial, int S, const float* __restrict__ bias, int relu,
                                    float scale, float* __restrict__ out, int ldo, int M, int cols) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * cols) return;
  const int m = i / cols, c = i - m * cols;
  float s = 0.f;
  for (int sp = 0; sp < S; ++sp) s += partial[((size_t)sp * M + m) * cols + c];
  s = s * scale + (bias ? bias[c] : 0.f);
  out[(size_t)m * ldo + c] = relu ? fmaxf(s, 0.f) : s;
}

// ---- TN: thread <-> 4 rows n x 4 consecutive k; warp = 128 consecutive k, 8 warps = 32 rows n -------------------------
__global__ void __launch_bounds__(256)
skinny_tn_kernel(const float* __restrict__ A, int lda, const float* __restrict__ ref, int ldref, const float* __restrict__ B,
                 int ldb, float* __restrict__ out, int ldo, float* __restrict__ out_bias, int M, int N, int K, int accumulate) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int k = (blockIdx.x * 32 + lane) * 4;
  const int n0 = (blockIdx.y * 8 + warp) * 4;
  if (n0 >= N) return;
  const bool k_ok = k < K;
  const bool full = n0 + 3 < N && (lda & 3) == 0 && (!ref || (ldref & 3) == 0);
  float acc[4][4], bsum[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int m = 0; m < M; ++m) {
    float a[4];
    if (full) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(A + (size_t)m * lda + n0));
      a[0] = v.x; a[1] = v.y; a[2] = v.z; a[3] = v.w;
      if (ref) {
        const float4 f = __ldg(reinterpret_cast<const float4*>(ref + (size_t)m * ldref + n0));
        if (!(f.x > 0.f)) a[0] = 0.f;
        if (!(f.y > 0.f)) a[1] = 0.f;
        if (!(f.z > 0.f)) a[2] = 0.f;
        if (!(f.w > 0.f)) a[3] = 0.f;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        a[i] = 0.f;
        if (n0 + i < N) {
          a[i] = A[(size_t)m * lda + n0 + i];
          if (ref && !(ref[(size_t)m * ldref + n0 + i] > 0.f)) a[i] = 0.f;
        }
      }
    }
    float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
    if (k_ok) b = __ldg(reinterpret_cast<const float4*>(B + (size_t)m * ldb + k));
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      acc[i][0] += a[i] * b.x; acc[i][1] += a[i] * b.y; acc[i][2] += a[i] * b.z; acc[i][3] += a[i] * b.w;
      bsum[i] += a[i];
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (n0 + i >= N) break;
    if (k_ok) {
      float4* dst = reinterpret_cast<float4*>(out + (size_t)(n0 + i) * ldo + k);
      float4 v = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
      if (accumulate) {                   // row blocks of a taller A are summed in call order
        const float4 p = *dst;
        v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w;
      }
      *dst = v;
    }
    if (out_bias && blockIdx.x == 0 && lane == 0) out_bias[n0 + i] = accumulate ? out_bias[n0 + i] + bsum[i] : bsum[i];
  }
}

}  // namespace b200

using namespace b200;

extern "C" size_t b200_skinny_gemm_workspace_bytes(int M, int cols) {
  return (size_t)kSkMaxSplit * max(M, 1) * max(cols, 1) * 4;
}

extern "C" int b200_skinny_gemm(int mode, const float* A, int lda, const float* relu_ref, int ldref, const float* B, int ldb,
                                const float* bias, int relu, float scale, float* out, int ldo, float* out_bias, int M,
                                int N, int K, int accumulate, void* workspace, size_t workspace_bytes,
                                b200_stream_t stream) {
  B200_CHECK_ARG(mode >= 0 && mode <= 2 && A && B && out, "skinny_gemm: bad mode or null tensor");
  B200_CHECK_ARG(M > 0 && M <= kSkMaxM && N > 0 && K > 0, "skinny_gemm: need 0 < M <= 32");
  B200_CHECK_ARG(!accumulate || mode == 2, "skinny_gemm: accumulate is a TN-mode option");
  B200_CHECK_ARG(!(relu_ref && mode == 0), "skinny_gemm: the ReLU mask applies to the NN / TN modes");
  // vector accesses: rows of B (all modes), rows of A along the reduction (NT: k, NN: n), rows of the TN output
  if (K % 4 || ldb % 4 || ((uintptr_t)B & 15) || (mode == 0 && (lda % 4 || ((uintptr_t)A & 15))) ||
      (mode == 1 && (N % 4 || lda % 4 || ((uintptr_t)A & 15) || (relu_ref && (ldref % 4 || ((uintptr_t)relu_ref & 15))))) ||
      (mode == 2 && (ldo % 4 || ((uintptr_t)out & 15) || ((uintptr_t)A & 15) || (relu_ref && ((uintptr_t)relu_ref & 15))))) {
    set_error("skinny_gemm: K (and N in NN mode) and the leading dimensions of the vector-accessed operands must be "
              "multiples of 4 floats, pointers 16-byte aligned");
    return B200_ERR_UNSUPPORTED;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (mode == 2) {
    dim3 grid(ceil_div(K, 128), ceil_div(N, 32));
    skinny_tn_kernel<<<grid, 256, 0, st>>>(A, lda, relu_ref, ldref, B, ldb, out, ldo, out_bias, M, N, K, accumulate);
    B200_CUDA_LAUNCH_CHECK("skinny_gemm(tn)");
    return B200_OK;
  }
  const int cols = mode == 0 ? N : K, reduce = mode == 0 ? K : N;
  B200_CHECK_ARG(workspace && workspace_bytes >= b200_skinny_gemm_workspace_bytes(M, cols), "skinny_gemm: workspace too small");
  float* ws = (float*)workspace;
  const int S = sk_split(reduce);
  dim3 grid(ceil_div(cols, kSkCols), S);
  if (mode == 0)
    skinny_reduce_kernel<0><<<grid, 256, 0, st>>>(A, lda, nullptr, 0, B, ldb, ws, M, cols, reduce);
  else
    skinny_reduce_kernel<1><<<grid, 256, 0, st>>>(A, lda, relu_ref, ldref, B, ldb, ws, M, cols, reduce);
  B200_CUDA_LAUNCH_CHECK("skinny_gemm(partial)");
  skinny_final_kernel<<<ceil_div(M * cols, 256), 256, 0, st>>>(ws, S, mode == 0 ? bias : nullptr, mode == 0 ? relu : 0,
                                                               mode == 0 ? 1.0f : scale, out, ldo, M, cols);
  B200_CUDA_LAUNCH_CHECK("skinny_gemm(final)");
  return B200_OK;
}
