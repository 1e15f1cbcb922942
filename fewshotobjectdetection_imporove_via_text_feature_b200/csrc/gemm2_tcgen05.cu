// CTA-pair bf16 tensor-core GEMM / implicit-GEMM convolution for sm_100a (tcgen05.mma.cta_group::2):
//   D[M,N] = epi( A[M,K] . B[N,K]^T )         bf16 operands, fp32 accumulation in TMEM
// used for the wide products of the ROI head: the res5 bottleneck convolutions and their data gradients
// (defrcn/modeling/roi_heads/roi_heads.py:313-344, detectron2 BottleneckBlock) and the large text-fusion GEMMs
// (attentive_modules.py:123-126,166-175,71-75) forward, dX and dW.
//
// A thread-block cluster of two CTAs (one SM pair) owns a 256 x BN output tile.  Per 64-wide K block each CTA brings its
// own 128 rows of A and its own BN/2 rows of B into shared memory with TMA (128B swizzle); the pair's leader issues ONE
// tcgen05.mma.cta_group::2 (M = 256, N = BN, K = 16) x 4 that reads both CTAs' shared memory, so every operand byte is
// fetched once per pair and each SM's shared-memory read traffic per flop is half that of a single-CTA tile.  Each CTA's
// TMEM holds the 128 x BN fp32 accumulator of its own rows, double-buffered: 8 epilogue warps per CTA drain tile i while
// the MMAs of tile i+1 run.  Persistent: cluster c walks tiles c, c + #clusters, ... (n fastest: the n-tiles of one
// 256-row block of A run back to back / side by side, so A streams from HBM once while the much smaller B — a weight
// matrix — stays L2-resident).
//
// A operand sources (per K block):
//   plain     2-D K-major [M][K] (TMA box 64 x 128), optionally a second tensor A2 for the tail of K ("K-concat":
//             out = [A | A2] . B^T — the bottleneck's conv3 + shortcut in one accumulator);
//   M-major   2-D [K][M] (the transposed operand of a weight-gradient product, no transposed copy in HBM);
//   conv3x3   4-D NHWC activation (R, 4, 4, C): K block kb = (tap, 64 channels); the box of 8 ROIs x 4 x 4 pixels x 64
//             channels is fetched at pixel offset (dy-1, dx-1) and TMA's out-of-bounds zero fill IS the padding.
// B: K-major [N][K] or N-major [K][N].
// Epilogue (fp32, per 32-column chunk, lane = output row): + bias[n], + residual[m][n] (bf16, TMA-loaded), ReLU,
// ReLU-backward gate from a packed 1-bit mask or a bf16 activation, then bf16 output through swizzled shared-memory
// staging and TMA stores (coalesced 64-byte rows, M / N tails clipped by TMA), an optional second bf16 copy, an optional
// fp32 output (direct 128-byte row stores, optionally accumulating), and the packed 1-bit mask (out > 0) for the backward.
#include <cuda.h>

#include "common.cuh"
#include "tcgen05.cuh"

namespace b200 {

constexpr int kG2BK = 64;            // 64 bf16 = 128 B = one swizzle-128B row
constexpr int kG2Threads = 320;      // warp 0: TMA producer, warp 1: TMEM alloc + MMA issue, warps 2..9: epilogue
constexpr int kG2EpiWarps = 8;
constexpr size_t kG2SplitCounterBytes = 65536;   // head of the split-K workspace: arrival counters (zero between launches)
constexpr int kG2BoxBytes = 4096;    // epilogue box: 32 rows x 64 bf16 (128-byte rows, 128B swizzle)
// epilogue features an instantiation carries (dead code for the others is compiled out)
constexpr uint32_t kFRes = 1, kFMaskBits = 2, kFMaskAct = 4, kFF32 = 8, kFBitsOut = 16, kFRowMean = 32, kFOut2 = 64;
constexpr uint32_t kFFwd = kFRes | kFBitsOut;                  // res5 forward: bias, residual, ReLU, mask out (packed bf16x2 tail)
constexpr uint32_t kFFwdMean = kFRes | kFBitsOut | kFRowMean;  // ... last block: + mean over the 16 pixels
constexpr uint32_t kFBwd = kFRes | kFMaskBits;                 // res5 backward: mask in, residual
constexpr uint32_t kFAll = 127;
// row-wise epilogues (one lane owns one output row, so these need no cross-lane traffic): sum of squares of a row's outputs,
// a per-row scale from such sums (cosine logits: attentive / my_module.py:449-469 as the epilogue of the two products), softmax
// over the row (attentive_modules.py:45-55 as the epilogue of the score product), gate operands O*x and x-O
// (attentive_modules.py:166,170 as the epilogue of the probabilities x values product).  Their own instantiation.
constexpr uint32_t kFRowSumSq = 128, kFRowScale = 256, kFSoftmax = 512, kFGate = 1024;
constexpr uint32_t kFRow = kFRes | kFF32 | kFOut2 | kFRowSumSq | kFRowScale | kFSoftmax | kFGate;

// packed fp32 add (sm_100: FADD2 adds two fp32 lanes of a 64-bit register pair per issue slot; the accumulator registers that
// tcgen05.ld fills are consecutive, so (r[2i], r[2i+1]) are natural pairs)
__device__ __forceinline__ void add_f32x2(uint32_t& a0, uint32_t& a1, uint32_t b0, uint32_t b1) {
  unsigned long long a, b;
  asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "r"(a0), "r"(a1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "r"(b0), "r"(b1));
  asm("add.rn.f32x2 %0, %0, %1;" : "+l"(a) : "l"(b));
  asm("mov.b64 {%0, %1}, %2;" : "=r"(a0), "=r"(a1) : "l"(a));
}

struct Gemm2Args {
  const float* bias;
  const __nv_bfloat16* mask_act;     // ReLU backward gate from a bf16 activation [M][ldmask] (zero where <= 0)
  int ldmask;
  const uint32_t* mask_bits;         // ... or from a packed mask [M][ldbits_in] (bit n % 32 of word n / 32)
  int ldbits_in;
  uint32_t* bits_out;                // packed mask of the output (value > 0 after the activation) [M][ldbits_out]
  int ldbits_out;
  float* d_f32;                      // optional fp32 output [M][ldd32] (direct stores), `accumulate`: D += result
  int ldd32;
  int accumulate;
  float* rowmean_out;                // optional: mean over each group of 16 consecutive rows -> [M/16][ld_rowmean] fp32
  int ld_rowmean;
  float* rowsumsq_out;               // optional: sum over each 64-column chunk of a row of the output's squares [M][ld_rowsumsq]
  int ld_rowsumsq;
  const float* row_sumsq_in;         // optional: per-row scale 1 / max(sqrt(sum of the row's row_parts entries), row_eps)
  int ld_row_sumsq_in, row_parts;
  float row_eps;
  int softmax;                       // out = softmax over the row's N <= 128 columns of (acc + bias)
  int gate;                          // residual operand is x: out = acc * x, out2 = x - acc
  int M, N;
  int kb1, kb2;                      // K blocks taken from A (map_a) and from A2 (map_a2)
  int conv_cb;                       // conv3x3 mode: K blocks per tap (C / 64); 0 = plain
  int a_mn, b_mn;                    // operand stored M- / N-contiguous
  int relu;
  int has_out, has_out2, has_res;    // map_d / map_d2 / map_res are valid
  int split_k;                       // > 1: the K blocks of a tile are shared by split_k work items (no residual then)
  float* ws;                         // split-K partial accumulators [split][Mp][Np] fp32
  int* counters;                     // split-K arrival counters [tile][rank][epilogue warp], zero between launches
  int ws_ld;                         // Np
  long long ws_split_stride;         // Mp * Np
};

template <int BN> struct G2Cfg {
  static constexpr int kABytes = 128 * kG2BK * 2;                 // 16 KB: this CTA's 128 rows of A
  static constexpr int kBBytes = (BN / 2) * kG2BK * 2;            // this CTA's half of the B tile
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = BN >= 256 ? 4 : 5;
  static constexpr int kTmemCols = 2 * BN;                        // two accumulators
  static constexpr int kEpiBytesPerWarp = 3 * kG2BoxBytes;         // output staging, residual x2
  static constexpr int kBarBytes = 512;
  static constexpr int kSmemBytes = kStages * kStageBytes + kG2EpiWarps * kEpiBytesPerWarp + kBarBytes + 1024 /*align*/;
};

template <int BN, uint32_t F>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kG2Threads, 1)
gemm2_pair_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_a2,
                  const __grid_constant__ CUtensorMap map_b, const __grid_constant__ CUtensorMap map_d,
                  const __grid_constant__ CUtensorMap map_d2, const __grid_constant__ CUtensorMap map_res,
                  const Gemm2Args p) {
  using Cfg = G2Cfg<BN>;
  constexpr int S = Cfg::kStages;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* tiles = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* epi = tiles + S * Cfg::kStageBytes;                                   // 1024-aligned
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(epi + kG2EpiWarps * Cfg::kEpiBytesPerWarp);
  uint64_t* empty_bar = full_bar + S;
  uint64_t* acc_full = empty_bar + S;
  uint64_t* acc_empty = acc_full + 2;
  uint64_t* res_bar = acc_empty + 2;                      // [warp][2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + 2 * kG2EpiWarps);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();                // 0 = leader of the pair
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  const int m_tiles = (p.M + 255) / 256, n_tiles = (p.N + BN - 1) / BN;
  const int num_tiles = m_tiles * n_tiles;
  const int num_kb = p.conv_cb ? 9 * p.conv_cb : p.kb1 + p.kb2;
  // split-K: work item w = (tile w / S, K slice w % S); the S slices of a tile run on S different CTA pairs at once
  const int S_k = p.split_k > 1 ? p.split_k : 1;
  const int kb_per = (num_kb + S_k - 1) / S_k;
  const int num_work = num_tiles * S_k;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    if (p.kb2) tma_prefetch_desc(&map_a2);
    if (p.has_out) tma_prefetch_desc(&map_d);
    if (p.has_out2) tma_prefetch_desc(&map_d2);
    if (p.has_res) tma_prefetch_desc(&map_res);
    for (int s = 0; s < S; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 2 * kG2EpiWarps); }
    for (int i = 0; i < 2 * kG2EpiWarps; ++i) mbar_init(&res_bar[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)Cfg::kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  cluster_sync_all();                                     // both CTAs' barriers are initialised, both TMEM allocations done
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // the next launch of the stream may take this SM as soon as this CTA leaves it and set itself up under our tail; our own
  // operands may still be in flight from the previous launch: nothing above touched global memory
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    // ---- TMA producer (both CTAs): own A rows, own half of B; the bytes of both CTAs complete on the LEADER's barrier
    if (lane == 0) {
      int it = 0;
      for (int w = cluster_id; w < num_work; w += num_clusters) {
        const int t = w / S_k, kb_lo = (w % S_k) * kb_per, kb_hi = min(num_kb, kb_lo + kb_per);
        const int m0 = (t / n_tiles) * 256 + (int)rank * 128;
        const int n0 = (t % n_tiles) * BN + (int)rank * (BN / 2);
        for (int kb = kb_lo; kb < kb_hi; ++kb, ++it) {
          const int s = it % S;
          const uint32_t ph = (it / S) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          const uint32_t sa = smem_u32(tiles + s * Cfg::kStageBytes);
          const uint32_t sb = sa + Cfg::kABytes;
          const uint32_t bar = mapa_u32(smem_u32(&full_bar[s]), 0);
          if (rank == 0) mbar_expect_tx(&full_bar[s], 2 * Cfg::kStageBytes);
          if (p.conv_cb) {
            const int tap = kb / p.conv_cb, kc = kb - tap * p.conv_cb;
            tma_load_4d_pair(sa, &map_a, bar, kc * kG2BK, tap % 3 - 1, tap / 3 - 1, m0 >> 4);
          } else if (p.a_mn) {
            const int k0 = kb * kG2BK;                   // source [K][M]: two boxes of 64 m x 64 k
            tma_load_2d_pair(sa, &map_a, bar, m0, k0);
            tma_load_2d_pair(sa + 8192, &map_a, bar, m0 + 64, k0);
          } else if (kb < p.kb1) {
            tma_load_2d_pair(sa, &map_a, bar, kb * kG2BK, m0);
          } else {
            tma_load_2d_pair(sa, &map_a2, bar, (kb - p.kb1) * kG2BK, m0);
          }
          if (p.b_mn) {
#pragma unroll
            for (int c = 0; c < BN / 128; ++c) tma_load_2d_pair(sb + c * 8192, &map_b, bar, n0 + 64 * c, kb * kG2BK);
          } else {
            tma_load_2d_pair(sb, &map_b, bar, kb * kG2BK, n0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ---- MMA issue: the leader's elected lane, for the pair
    if (rank == 0) {
      const uint32_t idesc = make_idesc_bf16(256, BN, p.a_mn, p.b_mn);
      // descriptor step per 16 k: K-major +32 B inside the swizzle row; MN-major +16 rows of 128 B
      const uint64_t a_step = p.a_mn ? (2048 >> 4) : (32 >> 4), b_step = p.b_mn ? (2048 >> 4) : (32 >> 4);
      const uint32_t a_lbo = p.a_mn ? 8192 : 0, b_lbo = p.b_mn ? 8192 : 0;
      int it = 0, lt = 0;
      for (int w = cluster_id; w < num_work; w += num_clusters, ++lt) {
        const int kb_lo = (w % S_k) * kb_per, kb_hi = min(num_kb, kb_lo + kb_per);
        const int buf = lt & 1;
        mbar_wait(&acc_empty[buf], ((lt >> 1) & 1) ^ 1);     // both CTAs' epilogues have drained this accumulator
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(buf * BN);
        for (int kb = kb_lo; kb < kb_hi; ++kb, ++it) {
          const int s = it % S;
          const uint32_t ph = (it / S) & 1;
          mbar_wait(&full_bar[s], ph);
          tcgen05_fence_after();
          if (elect_one()) {
            const uint32_t sa = smem_u32(tiles + s * Cfg::kStageBytes);
            const uint32_t sb = sa + Cfg::kABytes;
            const uint64_t adesc = make_smem_desc_sw128(sa, a_lbo), bdesc = make_smem_desc_sw128(sb, b_lbo);
#pragma unroll
            for (int k = 0; k < kG2BK / 16; ++k)
              umma_bf16_pair(tmem_d, adesc + a_step * k, bdesc + b_step * k, idesc, (kb > kb_lo) || (k > 0));
            umma_commit_pair(&empty_bar[s], 3);                        // both CTAs' stage s is reusable
            if (kb == kb_hi - 1) umma_commit_pair(&acc_full[buf], 3);  // both CTAs' accumulator halves are complete
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ---- epilogue: 8 warps per CTA; warp w owns TMEM lane quarter w % 4 (32 rows, one per lane) and one column half of
    // the tile, in 64-column chunks (two 32-column tcgen05.ld in flight together, 128-byte rows in the staging buffers).
    // F (compile time) lists the features this instantiation carries; the descriptor's flags choose among them.
    const int ew = warp - 2, q = warp & 3, half = ew >> 2;
    constexpr int kCw = BN / 2;                              // columns per warp
    constexpr int kChunks = kCw / 64;                        // 2 (BN = 256) or 1 (BN = 128)
    const int cbeg = half * kCw;
    unsigned char* my = epi + ew * Cfg::kEpiBytesPerWarp;
    const uint32_t s_out = smem_u32(my), s_res = s_out + 2 * kG2BoxBytes;     // output staging x2, residual box x1
    uint64_t* rbar = res_bar + 2 * ew;
    const int swz = lane & 7;                                // 128B swizzle: 16-byte chunk j of row r lives at j ^ (r & 7)
    const uint32_t row_off = (uint32_t)lane * 128;
    const bool f_res = (F & kFRes) && p.has_res;
    const bool f_mbits = (F & kFMaskBits) && p.mask_bits;
    const bool f_mact = (F & kFMaskAct) && p.mask_act;
    const bool f_f32 = (F & kFF32) && p.d_f32;
    const bool f_bout = (F & kFBitsOut) && p.bits_out;
    const bool f_mean = (F & kFRowMean) && p.rowmean_out;
    const bool f_out2 = (F & kFOut2) && p.has_out2;
    const bool f_store = p.has_out || f_out2;
    const bool f_ssq = (F & kFRowSumSq) && p.rowsumsq_out;
    const bool f_rscale = (F & kFRowScale) && p.row_sumsq_in;
    const bool f_softmax = (F & kFSoftmax) && p.softmax;
    const bool f_gate = (F & kFGate) && p.gate;
    const bool bias_vec = p.bias && (reinterpret_cast<uintptr_t>(p.bias) & 15) == 0;
    const bool vec32 = f_f32 && (p.ldd32 % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.d_f32) & 15) == 0);
    const bool vecm = f_mact && (p.ldmask % 8 == 0) && ((reinterpret_cast<uintptr_t>(p.mask_act) & 15) == 0);
    int lt = 0;
    uint32_t gc = 0;                                         // running chunk counter of this warp (residual buffer parity)
    auto issue_res = [&](int t, int ci, uint32_t g) {        // lane 0: residual box of tile t, chunk ci (chunk counter g)
      const int m0 = (t / n_tiles) * 256 + (int)rank * 128 + q * 32;
      const int n0 = (t % n_tiles) * BN + cbeg + ci * 64;
      mbar_expect_tx(&rbar[0], kG2BoxBytes);
      tma_load_2d_u32(s_res, &map_res, &rbar[0], n0, m0);
    };
    // one residual box per warp: the box of chunk g + 1 is requested as soon as chunk g's has been read into registers,
    // i.e. a whole chunk's processing ahead of its use, and it comes out of L2 (prefetched a tile ahead)
    if (f_res && lane == 0 && cluster_id < num_tiles) issue_res(cluster_id, 0, 0);
    for (int w = cluster_id; w < num_work; w += num_clusters, ++lt) {
      const int t = w / S_k;
      const int m0 = (t / n_tiles) * 256 + (int)rank * 128, n0 = (t % n_tiles) * BN;
      const int buf = lt & 1;
      const int row = m0 + q * 32 + lane;
      const bool row_ok = row < p.M;
      float rscale = 1.f;
      if (f_rscale && row_ok) {
        float ss = 0.f;
        for (int j = 0; j < p.row_parts; ++j) ss += __ldg(p.row_sumsq_in + (size_t)row * p.ld_row_sumsq_in + j);
        rscale = 1.f / fmaxf(sqrtf(ss), p.row_eps);
      }
      if (f_res && lane == 0 && t + num_clusters < num_tiles) {      // next tile's residual boxes -> L2
        const int tn = t + num_clusters;
        const int pm0 = (tn / n_tiles) * 256 + (int)rank * 128 + q * 32, pn0 = (tn % n_tiles) * BN + cbeg;
#pragma unroll
        for (int ci = 0; ci < kChunks; ++ci)
          asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(&map_res), "r"(pn0 + 64 * ci), "r"(pm0) : "memory");
      }
      // packed ReLU masks: this lane's row has kCw / 32 consecutive words per tile — read once, before the accumulator is
      // ready; the output's own mask words are collected and written once at the end of the tile
      uint32_t mw[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu}, bw[4] = {0u, 0u, 0u, 0u};
      if (f_mbits) {
#pragma unroll
        for (int c = 0; c < kCw / 32; ++c) {
          const int col = n0 + cbeg + 32 * c;
          mw[c] = (row_ok && col < p.N) ? __ldg(p.mask_bits + (size_t)row * p.ldbits_in + (col >> 5)) : 0u;
        }
      }
      mbar_wait(&acc_full[buf], (lt >> 1) & 1);
      tcgen05_fence_after();
      if (S_k > 1) {
        // split-K: every slice leaves its fp32 partial tile in the workspace; the slice that arrives LAST at the tile's
        // counter (per epilogue warp: the warps' regions are disjoint) sums the S partials in slice order 0..S-1 —
        // the same order whoever arrives last, so the result is bitwise reproducible — and runs the epilogue
        float* wrow = p.ws + (long long)(w % S_k) * p.ws_split_stride + (long long)row * p.ws_ld + n0 + cbeg;
#pragma unroll 1
        for (int ci = 0; ci < kChunks; ++ci) {
          uint32_t r[64];
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + cbeg + ci * 64);
          tmem_ld_32x32_nowait(taddr, r);
          tmem_ld_32x32_nowait(taddr + 32, r + 32);
          tmem_ld_wait();
          if (row_ok && n0 + cbeg + ci * 64 < p.N) {
            float4* dst = reinterpret_cast<float4*>(wrow + ci * 64);
#pragma unroll
            for (int i = 0; i < 16; ++i)
              __stcg(dst + i, make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]), __uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3])));
          }
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(&acc_empty[buf], 0);
        __threadfence();
        __syncwarp();
        int* cnt = p.counters + ((t * 2 + (int)rank) * kG2EpiWarps + ew);
        int prev = 0;
        if (lane == 0) prev = atomicAdd(cnt, 1);
        prev = __shfl_sync(0xffffffffu, prev, 0);
        if (prev != S_k - 1) continue;                       // another slice finishes this tile
        if (lane == 0) *cnt = 0;                             // ready for the next launch
        __threadfence();
      }
#pragma unroll 1
      for (int ci = 0; ci < kChunks; ++ci, ++gc) {
        const int c0 = cbeg + ci * 64;
        const int col0 = n0 + c0;
        uint32_t r[64];
        if (S_k > 1) {
#pragma unroll
          for (int i = 0; i < 64; ++i) r[i] = 0u;
          if (row_ok && col0 < p.N) {
            const float* src = p.ws + (long long)row * p.ws_ld + n0 + c0;
            for (int sl = 0; sl < S_k; ++sl, src += p.ws_split_stride) {
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const float4 v4 = __ldcg(reinterpret_cast<const float4*>(src) + i);
                r[4 * i] = __float_as_uint(__uint_as_float(r[4 * i]) + v4.x);
                r[4 * i + 1] = __float_as_uint(__uint_as_float(r[4 * i + 1]) + v4.y);
                r[4 * i + 2] = __float_as_uint(__uint_as_float(r[4 * i + 2]) + v4.z);
                r[4 * i + 3] = __float_as_uint(__uint_as_float(r[4 * i + 3]) + v4.w);
              }
            }
          }
        } else {
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + c0);
          tmem_ld_32x32_nowait(taddr, r);
          tmem_ld_32x32_nowait(taddr + 32, r + 32);
          tmem_ld_wait();
          if (ci == kChunks - 1) {                           // last read of this accumulator: hand it back to the MMA warp
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(&acc_empty[buf], 0);
          }
        }
        if (f_rscale) {
#pragma unroll
          for (int i = 0; i < 64; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * rscale);
        }
        if (f_gate) {
          // gate operands of the attention output O (this accumulator) and the query feature x (the residual operand's box):
          // out = O * x, out2 = x - O, both through the two staging boxes
          mbar_wait(&rbar[0], gc & 1);
          if (lane == 0) tma_store_wait_read<0>();
          __syncwarp();
          const uint32_t xbase = s_res + row_off;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            uint32_t w0, w1, w2, w3;
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3) : "r"(xbase + ((j ^ swz) << 4)));
            const uint32_t w[4] = {w0, w1, w2, w3};
            uint32_t g1[4], g2[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float x0 = __uint_as_float(w[k] << 16), x1 = __uint_as_float(w[k] & 0xffff0000u);
              const float o0 = __uint_as_float(r[8 * j + 2 * k]), o1 = __uint_as_float(r[8 * j + 2 * k + 1]);
              const __nv_bfloat162 a = __floats2bfloat162_rn(o0 * x0, o1 * x1), b = __floats2bfloat162_rn(x0 - o0, x1 - o1);
              g1[k] = *reinterpret_cast<const uint32_t*>(&a);
              g2[k] = *reinterpret_cast<const uint32_t*>(&b);
            }
            asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(s_out + row_off + ((j ^ swz) << 4)), "r"(g1[0]), "r"(g1[1]), "r"(g1[2]), "r"(g1[3]) : "memory");
            asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(s_out + kG2BoxBytes + row_off + ((j ^ swz) << 4)), "r"(g2[0]), "r"(g2[1]), "r"(g2[2]), "r"(g2[3]) : "memory");
          }
          fence_proxy_async_smem();
          __syncwarp();                                      // every lane has read its x row and written both boxes
          if (lane == 0) {
            if (ci + 1 < kChunks) issue_res(t, ci + 1, gc + 1);
            else if (t + num_clusters < num_tiles) issue_res(t + num_clusters, 0, gc + 1);
            if (col0 < p.N) {
              tma_store_2d(&map_d, s_out, col0, m0 + q * 32);
              tma_store_2d(&map_d2, s_out + kG2BoxBytes, col0, m0 + q * 32);
              tma_store_commit();
            }
          }
          continue;
        }
        if (f_softmax) {
          // softmax over the row: N <= 64 is local to this lane; 64 < N <= 128 (BN = 128: the other column half belongs to
          // warp ew ^ 4) merges the two halves' (max, sum) through this warp's otherwise unused residual box
          float mx = -3.0e38f;
#pragma unroll
          for (int i = 0; i < 64; ++i) {
            const bool ok = col0 + i < p.N;
            const float vv = ok ? __uint_as_float(r[i]) + (p.bias ? __ldg(p.bias + col0 + i) : 0.f) : -3.0e38f;
            r[i] = __float_as_uint(vv);
            mx = fmaxf(mx, vv);
          }
          float sum = 0.f;
#pragma unroll
          for (int i = 0; i < 64; ++i) {
            const float e = (col0 + i < p.N) ? __expf(__uint_as_float(r[i]) - mx) : 0.f;
            r[i] = __float_as_uint(e);
            sum += e;
          }
          float scale = 1.f / sum;
          if (p.N > 64) {
            float2* mine = reinterpret_cast<float2*>(my + 2 * kG2BoxBytes) + (lt & 1) * 32;
            const float2* theirs = reinterpret_cast<const float2*>(epi + (ew ^ 4) * Cfg::kEpiBytesPerWarp + 2 * kG2BoxBytes) + (lt & 1) * 32;
            mine[lane] = make_float2(mx, sum);
            asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
            const float2 o = theirs[lane];
            const float m2 = fmaxf(mx, o.x);
            const float tot = sum * __expf(mx - m2) + o.y * __expf(o.x - m2);
            scale = __expf(mx - m2) / tot;
          }
          if (f_store) {
            if (lane == 0) tma_store_wait_read<1>();
            __syncwarp();
          }
          const uint32_t sbase = s_out + (gc & 1) * kG2BoxBytes + row_off;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            uint32_t hw4[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float a0 = __uint_as_float(r[8 * j + 2 * k]) * scale, a1 = __uint_as_float(r[8 * j + 2 * k + 1]) * scale;
              r[8 * j + 2 * k] = __float_as_uint(a0);
              r[8 * j + 2 * k + 1] = __float_as_uint(a1);
              const __nv_bfloat162 hh = __floats2bfloat162_rn(a0, a1);
              hw4[k] = *reinterpret_cast<const uint32_t*>(&hh);
            }
            if (f_store)
              asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(sbase + ((j ^ swz) << 4)), "r"(hw4[0]), "r"(hw4[1]), "r"(hw4[2]), "r"(hw4[3]) : "memory");
          }
          if (f_f32 && row_ok) {
#pragma unroll
            for (int i = 0; i < 64; ++i)
              if (col0 + i < p.N) p.d_f32[(size_t)row * p.ldd32 + col0 + i] = __uint_as_float(r[i]);
          }
          if (f_store) {
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0 && col0 < p.N) {
              if (p.has_out) tma_store_2d(&map_d, s_out + (gc & 1) * kG2BoxBytes, col0, m0 + q * 32);
              tma_store_commit();
            }
          }
          continue;
        }
        if (f_res) {
          // residual (bf16, 128B-swizzled rows of the TMA box) added straight into the accumulator registers, so that the
          // box is free again — and the next one requested — before the rest of the chunk is processed
          mbar_wait(&rbar[0], gc & 1);
          const uint32_t base = s_res + row_off;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            uint32_t w0, w1, w2, w3;
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3) : "r"(base + ((j ^ swz) << 4)));
            const uint32_t w[4] = {w0, w1, w2, w3};
#pragma unroll
            for (int k = 0; k < 4; ++k) add_f32x2(r[8 * j + 2 * k], r[8 * j + 2 * k + 1], w[k] << 16, w[k] & 0xffff0000u);
          }
          __syncwarp();                                      // every lane has read its row: the box can be refilled
          if (lane == 0) {
            if (ci + 1 < kChunks) issue_res(t, ci + 1, gc + 1);
            else if (t + num_clusters < num_tiles) issue_res(t + num_clusters, 0, gc + 1);
          }
        }
        if (col0 >= p.N) continue;                           // warp-uniform: chunk entirely beyond N
        if (f_store) {                                       // staging buffer gc & 1 was last read by the store of chunk gc - 2
          if (lane == 0) {
            if (f_mean) tma_store_wait_read<0>();            // the mean's scratch is the other buffer
            else tma_store_wait_read<1>();
          }
          __syncwarp();
        }
        float ssq = 0.f;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int col = col0 + 32 * h;
          if (col >= p.N) continue;                          // warp-uniform
          const bool full = col + 32 <= p.N;
          float v[32];
          if (p.bias) {
            if (bias_vec && full) {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col) + i);
                uint32_t t0 = r[32 * h + 4 * i], t1 = r[32 * h + 4 * i + 1], t2 = r[32 * h + 4 * i + 2], t3 = r[32 * h + 4 * i + 3];
                add_f32x2(t0, t1, __float_as_uint(b4.x), __float_as_uint(b4.y));
                add_f32x2(t2, t3, __float_as_uint(b4.z), __float_as_uint(b4.w));
                v[4 * i] = __uint_as_float(t0); v[4 * i + 1] = __uint_as_float(t1);
                v[4 * i + 2] = __uint_as_float(t2); v[4 * i + 3] = __uint_as_float(t3);
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[32 * h + i]) + (col + i < p.N ? __ldg(p.bias + col + i) : 0.f);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[32 * h + i]);
          }
          // When nothing after the activation needs fp32 values (no gate, no fp32 / mean output), ReLU and the output's
          // mask run on the packed bf16x2 words instead: max.bf16x2 and set.ne.bf16x2 handle two columns per instruction
          // (ReLU commutes with the rounding; after it a stored value is positive iff it is not zero).
          constexpr bool packed_tail = (F & (kFMaskBits | kFMaskAct | kFF32 | kFRowMean)) == 0;   // such instantiations always store
          if (p.relu && !packed_tail) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
          }
          if (f_mbits) {      // packed mask word: bit j <-> column 2j, bit 16 + j <-> column 2j + 1
            const uint32_t pbits = (kChunks == 2 && ci) ? mw[2 + h] : mw[h];
#pragma unroll
            for (int i = 0; i < 32; ++i) if (!(pbits & (1u << ((i >> 1) | ((i & 1) << 4))))) v[i] = 0.f;
          }
          if (f_mact) {
            if (vecm && full && row_ok) {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const uint4 pm = __ldg(reinterpret_cast<const uint4*>(p.mask_act + (size_t)row * p.ldmask + col) + i);
                const uint32_t w[4] = {pm.x, pm.y, pm.z, pm.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  if ((w[j] & 0x8000u) || !(w[j] & 0x7fffu)) v[8 * i + 2 * j] = 0.f;
                  if ((w[j] & 0x80000000u) || !(w[j] & 0x7fff0000u)) v[8 * i + 2 * j + 1] = 0.f;
                }
              }
            } else if (row_ok) {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (col + i < p.N && !(__bfloat162float(p.mask_act[(size_t)row * p.ldmask + col + i]) > 0.f)) v[i] = 0.f;
            }
          }
          if (f_f32 && row_ok) {
            if (vec32 && full) {
              float4* dst = reinterpret_cast<float4*>(p.d_f32 + (size_t)row * p.ldd32 + col);
              if (p.accumulate) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const float4 pa = dst[i];
                  v[4 * i] += pa.x; v[4 * i + 1] += pa.y; v[4 * i + 2] += pa.z; v[4 * i + 3] += pa.w;
                }
              }
#pragma unroll
              for (int i = 0; i < 8; ++i) dst[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                if (col + i >= p.N) continue;
                float* dst = p.d_f32 + (size_t)row * p.ldd32 + col + i;
                if (p.accumulate) v[i] += *dst;
                *dst = v[i];
              }
            }
          }
          if (f_ssq) {
#pragma unroll
            for (int i = 0; i < 32; ++i) if (col + i < p.N) ssq += v[i] * v[i];
          }
          if (f_bout && !packed_tail) {
            uint32_t w = 0;
#pragma unroll
            for (int i = 0; i < 32; ++i) if (v[i] > 0.f) w |= 1u << ((i >> 1) | ((i & 1) << 4));
            if (kChunks == 2 && ci) bw[2 + h] = w; else bw[h] = w;
          }
          if (f_mean) {
            // mean over the 16 rows (the 4x4 pixels) of each ROI: lanes 0-15 hold one ROI, lanes 16-31 the next.  The warp's
            // 32 x 32 fp32 block goes through an output staging box (box h when nothing is stored; with a store, the box the
            // store of this chunk does not use — the chunk then starts by waiting for every earlier store's read), 16-byte
            // units XOR-swizzled by the row so that both the row writes and the column reads are conflict-free; lane l then
            // sums columns 2l', 2l'+1 (l' = l & 15) over its ROI's 16 rows in row order — 8 STS.128 + 16 LDS.64 + 32 FADD per
            // lane instead of a 30-shuffle butterfly.
            const uint32_t mbase = s_out + (f_store ? ((gc & 1) ^ 1) : h) * kG2BoxBytes;
            __syncwarp();                                     // the previous use of this box (chunk gc - 1) has been read
#pragma unroll
            for (int j = 0; j < 8; ++j)
              asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(mbase + row_off + ((j ^ swz) << 4)),
                           "f"(v[4 * j]), "f"(v[4 * j + 1]), "f"(v[4 * j + 2]), "f"(v[4 * j + 3]) : "memory");
            __syncwarp();
            const int l15 = lane & 15;
            const uint32_t rbase = mbase + (uint32_t)(lane & 16) * 128 + (uint32_t)(lane & 1) * 8;
            float s0 = 0.f, s1 = 0.f;
#pragma unroll
            for (int rr = 0; rr < 16; ++rr) {
              float a0, a1;
              asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(a0), "=f"(a1) : "r"(rbase + rr * 128 + ((((l15 >> 1) ^ (rr & 7))) << 4)));
              s0 += a0; s1 += a1;
            }
            const int roi = ((m0 + q * 32) >> 4) + (lane >> 4);
            const int cc = col + 2 * l15;
            if (roi * 16 < p.M) {
              float* dst = p.rowmean_out + (size_t)roi * p.ld_rowmean + cc;
              if (cc < p.N) dst[0] = s0 * (1.f / 16.f);
              if (cc + 1 < p.N) dst[1] = s1 * (1.f / 16.f);
            }
          }
          if (f_store) {
            const uint32_t base = s_out + (gc & 1) * kG2BoxBytes + row_off;
            uint32_t hw[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const __nv_bfloat162 hh = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
              hw[j] = *reinterpret_cast<const uint32_t*>(&hh);
            }
            if (packed_tail) {
              if (p.relu) {
#pragma unroll
                for (int j = 0; j < 16; ++j) asm("max.bf16x2 %0, %1, %2;" : "=r"(hw[j]) : "r"(hw[j]), "r"(0u));
              }
              if (f_bout) {
                uint32_t w = 0;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  uint32_t ne;                               // 0xffff per half that is not zero
                  asm("set.ne.u32.bf16x2 %0, %1, %2;" : "=r"(ne) : "r"(hw[j]), "r"(0u));
                  w |= ne & ((1u << j) | (1u << (16 + j)));
                }
                // without ReLU a stored value may be negative: positive = not zero and sign clear
                if (!p.relu) {
                  uint32_t neg = 0;
#pragma unroll
                  for (int j = 0; j < 16; ++j) neg |= ((hw[j] >> 15) & 1u) << j | ((hw[j] >> 31) & 1u) << (16 + j);
                  w &= ~neg;
                }
                if (kChunks == 2 && ci) bw[2 + h] = w; else bw[h] = w;
              }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j)
              asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(base + (((4 * h + j) ^ swz) << 4)),
                           "r"(hw[4 * j]), "r"(hw[4 * j + 1]), "r"(hw[4 * j + 2]), "r"(hw[4 * j + 3])
                           : "memory");
          }
        }
        if (f_ssq && row_ok) p.rowsumsq_out[(size_t)row * p.ld_rowsumsq + (col0 >> 6)] = ssq;
        if (f_store) {
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            if (p.has_out) tma_store_2d(&map_d, s_out + (gc & 1) * kG2BoxBytes, col0, m0 + q * 32);
            if (f_out2) tma_store_2d(&map_d2, s_out + (gc & 1) * kG2BoxBytes, col0, m0 + q * 32);
            tma_store_commit();
          }
        }
      }
      if (f_bout && row_ok) {
        uint32_t* dst = p.bits_out + (size_t)row * p.ldbits_out + ((n0 + cbeg) >> 5);
        if (kCw == 128 && n0 + cbeg + 128 <= p.N && (p.ldbits_out & 3) == 0 && (reinterpret_cast<uintptr_t>(p.bits_out) & 15) == 0) {
          *reinterpret_cast<uint4*>(dst) = make_uint4(bw[0], bw[1], bw[2], bw[3]);
        } else {
#pragma unroll
          for (int c = 0; c < kCw / 32; ++c) if (n0 + cbeg + 32 * c < p.N) dst[c] = bw[c];
        }
      }
    }
    if (lane == 0) tma_store_wait<0>();
  }
  tcgen05_fence_before();
  cluster_sync_all();            // the peer may still be reading this CTA's shared memory / signalling its barriers
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::kTmemCols) : "memory");
  }
}

// ---- host side --------------------------------------------------------------------------------------------------------
static int encode_map(CUtensorMap* m, const void* ptr, int rank, const cuuint64_t* dims, const cuuint64_t* strides,
                      const cuuint32_t* box, CUtensorMapSwizzle swz, const char* what) {
  PFN_encodeTiled enc = get_tensor_map_encoder();
  if (!enc) { set_error("gemm2: cuTensorMapEncodeTiled unavailable"); return B200_ERR_CUDA; }
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("gemm2: cuTensorMapEncodeTiled(%s) failed (%d)", what, (int)r); return B200_ERR_CUDA; }
  return B200_OK;
}
// 2-D bf16 [rows][cols] row-major, leading dimension ld (elements)
static int map_2d(CUtensorMap* m, const void* ptr, int rows, int cols, int ld, int box_cols, int box_rows,
                  CUtensorMapSwizzle swz, const char* what) {
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  return encode_map(m, ptr, 2, dims, strides, box, swz, what);
}

template <int BN, uint32_t F>
static int launch_gemm2(const b200_gemm2_desc* d, cudaStream_t st) {
  CUtensorMap ma, ma2, mb, md, md2, mr;
  Gemm2Args a = {};
  const int K1 = d->K, K2 = d->A2 ? d->K2 : 0;
  int rc;
  if (d->conv_c) {
    // A: (R, 4, 4, C) NHWC; box = 64 channels x 4 x 4 pixels x 8 ROIs = 128 rows of 128 B
    cuuint64_t dims[4] = {(cuuint64_t)d->conv_c, 4, 4, (cuuint64_t)(d->M / 16)};
    cuuint64_t strides[3] = {(cuuint64_t)d->conv_c * 2, (cuuint64_t)d->conv_c * 8, (cuuint64_t)d->conv_c * 32};
    cuuint32_t box[4] = {(cuuint32_t)kG2BK, 4, 4, 8};
    rc = encode_map(&ma, d->A, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B, "A conv");
  } else if (d->a_mn) {
    rc = map_2d(&ma, d->A, K1, d->M, d->lda, 64, kG2BK, CU_TENSOR_MAP_SWIZZLE_128B, "A mn");
  } else {
    rc = map_2d(&ma, d->A, d->M, K1, d->lda, kG2BK, 128, CU_TENSOR_MAP_SWIZZLE_128B, "A");
  }
  if (rc != B200_OK) return rc;
  ma2 = ma;
  if (K2) {
    rc = map_2d(&ma2, d->A2, d->M, K2, d->lda2, kG2BK, 128, CU_TENSOR_MAP_SWIZZLE_128B, "A2");
    if (rc != B200_OK) return rc;
  }
  const int Ktot = d->conv_c ? 9 * d->conv_c : K1 + K2;
  if (d->b_mn) rc = map_2d(&mb, d->B, Ktot, d->N, d->ldb, 64, kG2BK, CU_TENSOR_MAP_SWIZZLE_128B, "B mn");
  else rc = map_2d(&mb, d->B, d->N, Ktot, d->ldb, kG2BK, BN / 2, CU_TENSOR_MAP_SWIZZLE_128B, "B");
  if (rc != B200_OK) return rc;
  md = ma; md2 = ma; mr = ma;
  if (d->out_bf16) {
    rc = map_2d(&md, d->out_bf16, d->M, d->N, d->ld_out, 64, 32, CU_TENSOR_MAP_SWIZZLE_128B, "out");
    if (rc != B200_OK) return rc;
  }
  if (d->out2_bf16) {
    rc = map_2d(&md2, d->out2_bf16, d->M, d->N, d->ld_out2, 64, 32, CU_TENSOR_MAP_SWIZZLE_128B, "out2");
    if (rc != B200_OK) return rc;
  }
  if (d->residual) {
    rc = map_2d(&mr, d->residual, d->M, d->N, d->ld_res, 64, 32, CU_TENSOR_MAP_SWIZZLE_128B, "residual");
    if (rc != B200_OK) return rc;
  }
  a.bias = d->bias;
  a.mask_act = (const __nv_bfloat16*)d->mask_act; a.ldmask = d->ld_mask;
  a.mask_bits = (const uint32_t*)d->mask_bits; a.ldbits_in = d->ld_mask_bits;
  a.bits_out = (uint32_t*)d->bits_out; a.ldbits_out = d->ld_bits_out;
  a.d_f32 = d->out_f32; a.ldd32 = d->ld_out_f32; a.accumulate = d->accumulate;
  a.rowmean_out = d->rowmean_out; a.ld_rowmean = d->ld_rowmean;
  a.rowsumsq_out = d->rowsumsq_out; a.ld_rowsumsq = d->ld_rowsumsq;
  a.row_sumsq_in = d->row_scale_sumsq; a.ld_row_sumsq_in = d->ld_row_scale_sumsq; a.row_parts = d->row_scale_parts;
  a.row_eps = d->row_scale_eps;
  a.softmax = d->softmax; a.gate = d->gate;
  a.M = d->M; a.N = d->N;
  a.kb1 = ceil_div(K1, kG2BK); a.kb2 = ceil_div(K2, kG2BK);
  a.conv_cb = d->conv_c / kG2BK;
  a.a_mn = d->a_mn; a.b_mn = d->b_mn; a.relu = d->relu;
  a.has_out = d->out_bf16 != nullptr; a.has_out2 = d->out2_bf16 != nullptr; a.has_res = d->residual != nullptr;
  a.split_k = d->split_k > 1 ? d->split_k : 1;
  if (a.split_k > 1) {
    const int mt = ceil_div(d->M, 256), nt = ceil_div(d->N, BN);
    const size_t part = (size_t)a.split_k * mt * 256 * nt * BN * 4, need = kG2SplitCounterBytes + part;
    if ((size_t)mt * nt * 2 * kG2EpiWarps * 4 > kG2SplitCounterBytes) {
      set_error("gemm2: split-K is meant for products with few output tiles (%d x %d tiles)", mt, nt);
      return B200_ERR_UNSUPPORTED;
    }
    const int nkb = d->conv_c ? 9 * (d->conv_c / kG2BK) : a.kb1 + a.kb2;
    if (!d->splitk_workspace || d->splitk_workspace_bytes < need || (reinterpret_cast<uintptr_t>(d->splitk_workspace) & 15)) {
      set_error("gemm2: split-K needs a 16-byte aligned workspace of %zu bytes (b200_gemm2_splitk_workspace_bytes)", need);
      return B200_ERR_WORKSPACE;
    }
    if (d->residual || ceil_div(nkb, a.split_k) * (a.split_k - 1) >= nkb) {
      set_error("gemm2: split-K excludes a residual operand and needs every K slice non-empty (K blocks %d, split %d)", nkb, a.split_k);
      return B200_ERR_INVALID;
    }
    // counters first, at a fixed place: a workspace shared by launches of different shapes (one stream) keeps them zero
    a.counters = (int*)d->splitk_workspace;
    a.ws = (float*)((unsigned char*)d->splitk_workspace + kG2SplitCounterBytes);
    a.ws_ld = nt * BN;
    a.ws_split_stride = (long long)mt * 256 * nt * BN;
  }

  auto kern = gemm2_pair_kernel<BN, F>;
  B200_CUDA_CALL(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, G2Cfg<BN>::kSmemBytes));
  const int tiles = ceil_div(d->M, 256) * ceil_div(d->N, BN) * a.split_k;
  int clusters = min(tiles, kNumSMs / 2);
  if (d->max_clusters > 0) clusters = min(clusters, d->max_clusters);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * clusters);
  cfg.blockDim = dim3(kG2Threads);
  cfg.dynamicSmemBytes = G2Cfg<BN>::kSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  if (!d->no_pdl) {
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
  }
  B200_CUDA_CALL(cudaLaunchKernelEx(&cfg, kern, ma, ma2, mb, md, md2, mr, a));
  return B200_OK;
}

}  // namespace b200

using namespace b200;

extern "C" size_t b200_gemm2_splitk_workspace_bytes(int M, int N, int tile_n, int split_k) {
  if (split_k <= 1 || M <= 0 || N <= 0 || (tile_n != 128 && tile_n != 256)) return 0;
  const size_t mt = (size_t)ceil_div(M, 256), nt = (size_t)ceil_div(N, tile_n);
  return kG2SplitCounterBytes + (size_t)split_k * mt * 256 * nt * tile_n * 4;
}

extern "C" int b200_gemm2(const b200_gemm2_desc* d, b200_stream_t stream) {
  B200_CHECK_ARG(d, "gemm2: null descriptor");
  B200_CHECK_ARG(d->A && d->B, "gemm2: null operand");
  B200_CHECK_ARG(d->out_bf16 || d->out2_bf16 || d->out_f32 || d->rowmean_out, "gemm2: no output");
  const bool row_ops = d->rowsumsq_out || d->row_scale_sumsq || d->softmax || d->gate;
  if (row_ops) {
    B200_CHECK_ARG(!d->mask_act && !d->mask_bits && !d->bits_out && !d->rowmean_out && d->split_k <= 1 && !d->accumulate,
                   "gemm2: row-wise epilogues exclude gates, mask / mean outputs, split-K and accumulate");
    B200_CHECK_ARG(!d->row_scale_sumsq || (d->row_scale_parts > 0 && d->ld_row_scale_sumsq >= d->row_scale_parts),
                   "gemm2: row_scale_sumsq needs row_scale_parts > 0 entries per row");
    B200_CHECK_ARG(!d->softmax || (d->N <= 128 && !d->residual && !d->relu && !d->out2_bf16 && !d->rowsumsq_out),
                   "gemm2: softmax epilogue needs N <= 128 and excludes residual / ReLU / out2 / rowsumsq");
    B200_CHECK_ARG(!d->gate || (d->residual && d->out_bf16 && d->out2_bf16 && !d->bias && !d->relu && !d->out_f32 && !d->softmax &&
                                !d->rowsumsq_out),
                   "gemm2: gate epilogue takes x as the residual operand and writes out (O*x) and out2 (x-O) only");
  }
  B200_CHECK_ARG(d->M >= 0 && d->N > 0 && d->K > 0, "gemm2: bad shape");
  B200_CHECK_ARG(!d->accumulate || d->out_f32, "gemm2: accumulate needs an fp32 output");
  B200_CHECK_ARG(!(d->mask_act && d->mask_bits), "gemm2: one ReLU-backward gate at most");
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (!al16(d->A) || !al16(d->B) || !al16(d->A2) || !al16(d->out_bf16) || !al16(d->out2_bf16) || !al16(d->residual)) {
    set_error("gemm2: bf16 tensors must be 16-byte aligned (TMA)");
    return B200_ERR_UNSUPPORTED;
  }
  if (d->lda % 8 || d->ldb % 8 || (d->A2 && d->lda2 % 8) || (d->out_bf16 && d->ld_out % 8) || (d->out2_bf16 && d->ld_out2 % 8) ||
      (d->residual && d->ld_res % 8)) {
    set_error("gemm2: leading dimensions of bf16 tensors must be multiples of 8 (TMA)");
    return B200_ERR_UNSUPPORTED;
  }
  if (d->conv_c) {
    B200_CHECK_ARG(d->conv_c % kG2BK == 0 && d->M % 16 == 0 && !d->A2 && !d->a_mn && d->K == 9 * d->conv_c,
                   "gemm2: conv3x3 mode needs C %% 64 == 0, M = 16 R, K = 9 C, a K-major 4x4 NHWC activation");
  } else {
    // K itself is free (TMA zero-fills the tail of the last 64-wide K block on both operands); only the row pitches of the
    // bf16 tensors are constrained (multiples of 8 elements, checked above)
    B200_CHECK_ARG(!d->A2 || (d->K % kG2BK == 0 && d->K2 > 0 && !d->a_mn && !d->b_mn),
                   "gemm2: K-concat needs K %% 64 == 0 and K-major operands");
  }
  B200_CHECK_ARG(!d->bits_out || (d->N % 32 == 0), "gemm2: bits_out needs N %% 32 == 0");
  B200_CHECK_ARG(!d->rowmean_out || d->M % 16 == 0, "gemm2: rowmean_out needs M %% 16 == 0");
  if (d->M == 0) return B200_OK;
  cudaStream_t st = (cudaStream_t)stream;
  int bn = d->tile_n;
  if (bn == 0) {
    // A 128-wide tile moves 3/4 of a 256-wide tile's operand bytes for half of its flops (the mainloop is L2-bound), so
    // 256 wins unless the 128-wide tiling still fits one wave of CTA pairs (or N itself is narrow)
    const int mt = ceil_div(d->M, 256);
    bn = (d->N <= 128 || mt * ceil_div(d->N, 128) <= kNumSMs / 2) ? 128 : 256;
  }
  if (d->softmax) bn = 128;                               // the row's two 64-column halves sit in the two warps of a lane quarter
  B200_CHECK_ARG(bn == 128 || bn == 256, "gemm2: tile_n must be 0, 128 or 256");
  B200_CHECK_ARG(d->split_k <= 1 || d->tile_n == bn, "gemm2: split-K needs an explicit tile_n (the workspace is sized for it)");
  // smallest instantiation whose compiled-in epilogue features cover the descriptor
  uint32_t need = 0;
  if (d->residual) need |= kFRes;
  if (d->mask_bits) need |= kFMaskBits;
  if (d->mask_act) need |= kFMaskAct;
  if (d->out_f32) need |= kFF32;
  if (d->bits_out) need |= kFBitsOut;
  if (d->rowmean_out) need |= kFRowMean;
  if (d->out2_bf16) need |= kFOut2;
  if (row_ops) return bn == 256 ? launch_gemm2<256, kFRow>(d, st) : launch_gemm2<128, kFRow>(d, st);
  if (d->epilogue_variant == 1) need = kFAll;          // tests: force the generic instantiation
  if ((need & ~kFFwd) == 0) return bn == 256 ? launch_gemm2<256, kFFwd>(d, st) : launch_gemm2<128, kFFwd>(d, st);
  if ((need & ~kFFwdMean) == 0) return bn == 256 ? launch_gemm2<256, kFFwdMean>(d, st) : launch_gemm2<128, kFFwdMean>(d, st);
  if ((need & ~kFBwd) == 0) return bn == 256 ? launch_gemm2<256, kFBwd>(d, st) : launch_gemm2<128, kFBwd>(d, st);
  return bn == 256 ? launch_gemm2<256, kFAll>(d, st) : launch_gemm2<128, kFAll>(d, st);
}
