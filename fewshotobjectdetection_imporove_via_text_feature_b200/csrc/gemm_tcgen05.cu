// bf16 tensor-core GEMM for the text-fusion chain (A1, A4, A5, C1, C1'), hand-written for sm_100a:
//   D[M,N] = act(A[M,K] . B[N,K]^T + bias[N])          A, B bf16 K-contiguous ("K-major"), fp32 accumulate
// Reference ops replaced: nn.Linear at defrcn/modeling/roi_heads/attentive_modules.py:124 (w_q),
// :166-175 (linear1/2/3), :72 (FFN linear1/2), fast_rcnn.py:407,415 (bbox_pred, cls_score),
// roi_heads.py:1157-1159 (output_projection and the product with the text prototypes).
//
// Structure (persistent: one CTA per SM walks 128 x BN output tiles, 320 threads):
//   warp 0      TMA producer: cp.async.bulk.tensor 2-D boxes of A (128 x 64) and B (BN x 64) bf16 into a
//               STAGES-deep ring of 128B-swizzled shared-memory tiles, completion on `full` mbarriers;
//   warp 1      allocates TMEM (two accumulators of BN fp32 columns x 128 lanes), then one elected lane issues
//               tcgen05.mma.cta_group::1.kind::f16 (M=128, N=BN, K=16) x4 per stage and releases the stage
//               with tcgen05.commit -> `empty` mbarrier; a final commit per tile signals `acc_full[buf]`;
//   warps 2..9  epilogue of tile i while the MMA warp already runs tile i+1 into the other accumulator:
//               tcgen05.ld 32x32b.x32 (warp w owns TMEM lane quarter w % 4 = 32 output rows and one column half),
//               `acc_empty[buf]` arrive after the last read, bias / ReLU / ReLU-backward mask / fp32 accumulate with
//               the global operands prefetched a chunk ahead, fp32 and/or bf16 16-byte stores.
// OOB handling is TMA's: boxes hanging over M, N or K are zero-filled, stores are masked.
#include <cuda.h>

#include "common.cuh"
#include "tcgen05.cuh"

namespace b200 {

constexpr int kBM = 128;
constexpr int kBK = 64;        // 64 bf16 = 128 B = one swizzle-128B row
constexpr int kUmmaK = 16;
constexpr int kGemmThreads = 320;   // TMA warp, MMA warp, 8 epilogue warps

constexpr int kEpiWarps = 8;
constexpr int kEpiStageBytes = 128 * 4;          // per epilogue warp: its bias slice (<= 128 columns)

template <int BN> struct GemmCfg {
  static constexpr int kStageBytes = (kBM + BN) * kBK * 2;
  static constexpr int kStages = (BN >= 256) ? 4 : (BN >= 128 ? 6 : 8);
  static constexpr int kAccBufs = 2;             // TMEM accumulators: tile i's epilogue overlaps tile i+1's MMAs
  static constexpr int kTmemCols = (kAccBufs * BN) < 32 ? 32 : kAccBufs * BN;   // power of two >= 32
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/ + kEpiWarps * kEpiStageBytes;
};

// Persistent: CTA b runs tiles b, b + gridDim.x, ... (m fastest, so that concurrently running CTAs share a B panel in L2).
template <int BN>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                         const float* __restrict__ bias, float* __restrict__ d_f32, __nv_bfloat16* __restrict__ d_bf16,
                         int ldd, __nv_bfloat16* __restrict__ d2, int ldd2, int M, int N, int K, int relu,
                         int accumulate, const __nv_bfloat16* __restrict__ mask, int ldmask, int vec) {
  using Cfg = GemmCfg<BN>;
  constexpr int S = Cfg::kStages;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* tiles = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tiles + S * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + S;
  uint64_t* acc_full = empty_bar + S;
  uint64_t* acc_empty = acc_full + Cfg::kAccBufs;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + Cfg::kAccBufs);
  float* s_epi = reinterpret_cast<float*>(tiles + S * Cfg::kStageBytes + 256);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = (M + kBM - 1) / kBM, n_tiles = (N + BN - 1) / BN;
  const int num_tiles = m_tiles * n_tiles;
  const int num_kb = (K + kBK - 1) / kBK;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    for (int s = 0; s < S; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < Cfg::kAccBufs; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], kEpiWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)Cfg::kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int m0 = (t % m_tiles) * kBM, n0 = (t / m_tiles) * BN;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % S;
          const uint32_t ph = (it / S) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          unsigned char* sa = tiles + s * Cfg::kStageBytes;
          unsigned char* sb = sa + kBM * kBK * 2;
          mbar_expect_tx(&full_bar[s], Cfg::kStageBytes);
          tma_load_2d(sa, &map_a, &full_bar[s], kb * kBK, m0);
          tma_load_2d(sb, &map_b, &full_bar[s], kb * kBK, n0);
        }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_bf16(kBM, BN);
    int it = 0, lt = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++lt) {
      const int buf = lt & 1;
      mbar_wait(&acc_empty[buf], ((lt >> 1) & 1) ^ 1);     // the epilogue has drained this accumulator
      tcgen05_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(buf * BN);
      for (int kb = 0; kb < num_kb; ++kb, ++it) {
        const int s = it % S;
        const uint32_t ph = (it / S) & 1;
        mbar_wait(&full_bar[s], ph);
        tcgen05_fence_after();
        if (elect_one()) {
          const uint32_t sa = smem_u32(tiles + s * Cfg::kStageBytes);
          const uint32_t sb = sa + kBM * kBK * 2;
          const uint64_t adesc = make_smem_desc_sw128(sa), bdesc = make_smem_desc_sw128(sb);
#pragma unroll
          for (int k = 0; k < kBK / kUmmaK; ++k) {
            // advance 16 bf16 = 32 B inside the 128 B swizzle row: +2 in 16-byte descriptor units
            umma_bf16(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0);
          }
          umma_commit(&empty_bar[s]);                         // stage reusable once these MMAs have read it
          if (kb == num_kb - 1) umma_commit(&acc_full[buf]);  // accumulator complete
        }
        __syncwarp();
      }
    }
  } else {
    // Epilogue: 8 warps.  Warp w reads TMEM lane quarter w % 4 (32 output rows, one per lane) and the column half
    // (w - 2) / 4 of the tile, in 32-column chunks.  A chunk's global operands (ReLU-backward mask, fp32 addend) are
    // fetched one chunk ahead, all of them at once, so their latency hides under the previous chunk's stores; the
    // tile's bias slice sits in a per-warp shared copy.  Running a whole tile behind the MMA warp (second TMEM
    // accumulator), the epilogue only has to keep up with the mainloop, not be fast.
    const int ew = warp - 2, q = warp & 3, half = ew >> 2;
    constexpr int kCw = BN >= 64 ? BN / 2 : BN;              // columns per warp; BN = 32: the second half has no work
    const bool has_cols = BN >= 64 || half == 0;
    const int cbeg = half * kCw;
    float* sb = s_epi + ew * 128;                            // kCw <= 128 floats
    int lt = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++lt) {
      const int m0 = (t % m_tiles) * kBM, n0 = (t / m_tiles) * BN;
      const int buf = lt & 1;
      const int row = m0 + q * 32 + lane;
      const bool row_ok = row < M;
      uint4 pm[4];
      float4 pa[8];
      auto prefetch = [&](int c0n) {
        const int col = n0 + c0n;
        const bool ok = vec && row_ok && c0n < cbeg + kCw && col + 32 <= N;
#pragma unroll
        for (int i = 0; i < 4; ++i)
          pm[i] = (mask && ok) ? __ldg(reinterpret_cast<const uint4*>(mask + (size_t)row * ldmask + col) + i)
                               : make_uint4(0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u);
#pragma unroll
        for (int i = 0; i < 8; ++i)
          pa[i] = (accumulate && ok) ? *(reinterpret_cast<const float4*>(d_f32 + (size_t)row * ldd + col) + i)
                                     : make_float4(0.f, 0.f, 0.f, 0.f);
      };
      if (has_cols) {
        __syncwarp();                                        // previous tile's reads of `sb` are done
        for (int i = lane; i < kCw; i += 32) sb[i] = (bias && n0 + cbeg + i < N) ? bias[n0 + cbeg + i] : 0.f;
        __syncwarp();
        prefetch(cbeg);
      }
      mbar_wait(&acc_full[buf], (lt >> 1) & 1);
      tcgen05_fence_after();
      if (!has_cols) {
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[buf]);
        continue;
      }
#pragma unroll 1
      for (int c0 = cbeg; c0 < cbeg + kCw; c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + c0), r);
        if (c0 + 32 >= cbeg + kCw) {                         // last read of this accumulator: hand it back to the MMA warp
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[buf]);
        }
        const int col = n0 + c0;
        if (col >= N) continue;                              // warp-uniform
        float v[32];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 b4 = *reinterpret_cast<const float4*>(sb + (c0 - cbeg) + 4 * i);
          v[4 * i] = __uint_as_float(r[4 * i]) + b4.x;
          v[4 * i + 1] = __uint_as_float(r[4 * i + 1]) + b4.y;
          v[4 * i + 2] = __uint_as_float(r[4 * i + 2]) + b4.z;
          v[4 * i + 3] = __uint_as_float(r[4 * i + 3]) + b4.w;
        }
        if (relu) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
        }
        if (vec && col + 32 <= N) {
          // ReLU backward: zero where the forward activation (bf16) was not positive: sign set or magnitude zero
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const uint32_t w[4] = {pm[i].x, pm[i].y, pm[i].z, pm[i].w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if ((w[j] & 0x8000u) || !(w[j] & 0x7fffu)) v[8 * i + 2 * j] = 0.f;
              if ((w[j] & 0x80000000u) || !(w[j] & 0x7fff0000u)) v[8 * i + 2 * j + 1] = 0.f;
            }
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) { v[4 * i] += pa[i].x; v[4 * i + 1] += pa[i].y; v[4 * i + 2] += pa[i].z; v[4 * i + 3] += pa[i].w; }
          prefetch(c0 + 32);                                 // next chunk's operands, in flight during the stores below
          if (row_ok) {
            if (d_f32) {
              float4* dst = reinterpret_cast<float4*>(d_f32 + (size_t)row * ldd + col);
#pragma unroll
              for (int i = 0; i < 8; ++i) dst[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
            }
            if (d_bf16 || d2) {
              uint4 w[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const __nv_bfloat162 h0 = __floats2bfloat162_rn(v[8 * i], v[8 * i + 1]), h1 = __floats2bfloat162_rn(v[8 * i + 2], v[8 * i + 3]);
                const __nv_bfloat162 h2 = __floats2bfloat162_rn(v[8 * i + 4], v[8 * i + 5]), h3 = __floats2bfloat162_rn(v[8 * i + 6], v[8 * i + 7]);
                w[i].x = *reinterpret_cast<const uint32_t*>(&h0); w[i].y = *reinterpret_cast<const uint32_t*>(&h1);
                w[i].z = *reinterpret_cast<const uint32_t*>(&h2); w[i].w = *reinterpret_cast<const uint32_t*>(&h3);
              }
              if (d_bf16) {
                uint4* dst = reinterpret_cast<uint4*>(d_bf16 + (size_t)row * ldd + col);
#pragma unroll
                for (int i = 0; i < 4; ++i) dst[i] = w[i];
              }
              if (d2) {
                uint4* dst = reinterpret_cast<uint4*>(d2 + (size_t)row * ldd2 + col);
#pragma unroll
                for (int i = 0; i < 4; ++i) dst[i] = w[i];
              }
            }
          }
        } else if (row_ok) {
          // element-wise tail: the chunk hangs over N, or an operand is not aligned for vector access
#pragma unroll
          for (int i = 0; i < 32; ++i) {                     // static indexing keeps v[] in registers
            if (col + i >= N) continue;
            float x = v[i];
            if (mask && !(__bfloat162float(mask[(size_t)row * ldmask + col + i]) > 0.f)) x = 0.f;
            if (d_f32) {
              float* dst = d_f32 + (size_t)row * ldd + col + i;
              if (accumulate) x += *dst;
              *dst = x;
            }
            const __nv_bfloat16 h = __float2bfloat16_rn(x);
            if (d_bf16) d_bf16[(size_t)row * ldd + col + i] = h;
            if (d2) d2[(size_t)row * ldd2 + col + i] = h;
          }
          prefetch(c0 + 32);
        }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::kTmemCols) : "memory");
  }
}

// ---- host side -------------------------------------------------------------------------------------
PFN_encodeTiled get_tensor_map_encoder() {
  static PFN_encodeTiled fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      return (PFN_encodeTiled)p;
    return (PFN_encodeTiled) nullptr;
  }();          // thread-safe one-time initialisation (C++11 static)
  return fn;
}

// 2-D bf16 row-major tensor [rows][cols] with leading dimension ld (elements); box = box_rows x 64 cols
static int make_map(CUtensorMap* m, const void* ptr, int rows, int cols, int ld, int box_rows) {
  PFN_encodeTiled enc = get_tensor_map_encoder();
  if (!enc) { set_error("gemm_bf16: cuTensorMapEncodeTiled unavailable"); return B200_ERR_CUDA; }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("gemm_bf16: cuTensorMapEncodeTiled failed (%d)", (int)r); return B200_ERR_CUDA; }
  return B200_OK;
}

int g_gemm_generic_epilogue = 0;   // tests: force the element-wise epilogue on shapes the vector one covers
int g_gemm_ctas = 0;   // 0: one persistent CTA per SM; tests lower it to force several tiles per CTA on small shapes

template <int BN>
static int launch_gemm(const void* A, int lda, const void* B, int ldb, const float* bias, void* D, int ldd, int out_dtype,
                       void* D2, int ldd2, int M, int N, int K, int relu, int accumulate, const void* mask, int ldmask,
                       cudaStream_t st) {
  CUtensorMap ma, mb;
  int rc = make_map(&ma, A, M, K, lda, kBM);
  if (rc != B200_OK) return rc;
  rc = make_map(&mb, B, N, K, ldb, BN);
  if (rc != B200_OK) return rc;
  auto kern = gemm_bf16_tcgen05_kernel<BN>;
  // per device and cheap: set unconditionally (a process-wide flag would miss a second GPU and race between threads)
  B200_CUDA_CALL(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<BN>::kSmemBytes));
  // vector epilogue (16-byte accesses on full 32-column chunks) needs aligned operand rows; chunks hanging over N and
  // unaligned operands take the element-wise path
  const uintptr_t al = (uintptr_t)D | (uintptr_t)D2 | (uintptr_t)mask;
  const int vec = (al & 15) == 0 && (!D || ldd % (out_dtype == B200_F32 ? 4 : 8) == 0) && (!D2 || ldd2 % 8 == 0) &&
                  (!mask || ldmask % 8 == 0) && !g_gemm_generic_epilogue;
  const int tiles = ceil_div(N, BN) * ceil_div(M, kBM);
  dim3 grid(min(tiles, g_gemm_ctas > 0 ? g_gemm_ctas : kNumSMs));
  kern<<<grid, kGemmThreads, GemmCfg<BN>::kSmemBytes, st>>>(ma, mb, bias, out_dtype == B200_F32 ? (float*)D : nullptr,
                                                            out_dtype == B200_BF16 ? (__nv_bfloat16*)D : nullptr, ldd,
                                                            (__nv_bfloat16*)D2, ldd2, M, N, K, relu, accumulate,
                                                            (const __nv_bfloat16*)mask, ldmask, vec);
  B200_CUDA_LAUNCH_CHECK("gemm_bf16");
  return B200_OK;
}

}  // namespace b200

using namespace b200;

extern "C" int b200_gemm_bf16_ex(const void* A, int lda, const void* B, int ldb, const float* bias, void* D, int ldd,
                                 int out_dtype, void* D2, int ldd2, int M, int N, int K, int relu, int accumulate,
                                 const void* mask, int ldmask, b200_stream_t stream) {
  B200_CHECK_ARG(A && B && (D || D2), "gemm_bf16: null tensor");
  B200_CHECK_ARG(M >= 0 && N > 0 && K > 0, "gemm_bf16: bad shape");
  B200_CHECK_ARG((out_dtype | 1) == 1, "gemm_bf16: bad out_dtype");
  B200_CHECK_ARG(!accumulate || (D && out_dtype == B200_F32), "gemm_bf16: accumulate needs an fp32 D");
  if (K % 8 || lda % 8 || ldb % 8 || ((uintptr_t)A & 15) || ((uintptr_t)B & 15)) {
    set_error("gemm_bf16: K, lda, ldb must be multiples of 8 and A, B 16-byte aligned (TMA)");
    return B200_ERR_UNSUPPORTED;
  }
  if (M == 0) return B200_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int mt = ceil_div(M, kBM);
  // widest tile that still gives ~one CTA per SM; narrow N gets the narrowest tile that covers it
  int bn;
  if (N <= 32) bn = 32;
  else if (N <= 64) bn = 64;
  else if (mt * ceil_div(N, 256) >= 120) bn = 256;
  else if (mt * ceil_div(N, 128) >= 100 || N <= 128) bn = 128;
  else if (mt * ceil_div(N, 64) >= kNumSMs / 2) bn = 64;
  else bn = 32;   // few tiles (the M <= 128 weight-gradient products, K = R): twice the CTAs streaming the K panel
  switch (bn) {
    case 32: return launch_gemm<32>(A, lda, B, ldb, bias, D, ldd, out_dtype, D2, ldd2, M, N, K, relu, accumulate, mask, ldmask, st);
    case 64: return launch_gemm<64>(A, lda, B, ldb, bias, D, ldd, out_dtype, D2, ldd2, M, N, K, relu, accumulate, mask, ldmask, st);
    case 128: return launch_gemm<128>(A, lda, B, ldb, bias, D, ldd, out_dtype, D2, ldd2, M, N, K, relu, accumulate, mask, ldmask, st);
    default: return launch_gemm<256>(A, lda, B, ldb, bias, D, ldd, out_dtype, D2, ldd2, M, N, K, relu, accumulate, mask, ldmask, st);
  }
}

extern "C" int b200_gemm_bf16(const void* A, int lda, const void* B, int ldb, const float* bias, void* D, int ldd,
                              int out_dtype, void* D2, int ldd2, int M, int N, int K, int relu, b200_stream_t stream) {
  return b200_gemm_bf16_ex(A, lda, B, ldb, bias, D, ldd, out_dtype, D2, ldd2, M, N, K, relu, 0, nullptr, 0, stream);
}
