// D1 + D2 + D3: fast_rcnn_inference on the device, no host synchronisation.
//   reference: defrcn/modeling/roi_heads/fast_rcnn.py:46-134 (+ :306-334 predict_boxes / predict_probs),
//   third-party arithmetic: detectron2 0.3 Box2BoxTransform.apply_deltas, Boxes.clip, layers.batched_nms,
//   torchvision 0.8.1 ops.nms (coordinate-offset trick).
//
// Kernels
//   softmax_decode_compact_kernel : one CTA per image.  Softmax over K+1 logits (warp per row), score
//       threshold, ORDERED compaction (row-major == torch.nonzero order, via ballot prefix + block scan),
//       box decode + clip fused into the compaction: only surviving (roi,class) pairs are decoded.
//   nms_prepare_kernel            : one CTA per image.  max coordinate (for the offset trick), class
//       histogram, stable counting sort of candidates by class.
//   nms_class_kernel              : one CTA per (class, image).  shared-memory bitonic sort by
//       (score desc, index asc); 64-box diagonal blocks resolved with an IoU bitmask, then a parallel
//       sweep of the survivors over the remaining boxes with a shared-memory `removed` bitmap.
//   nms_finalize_kernel           : one CTA per image.  Merge per-class survivors by (score desc, idx asc),
//       emit the first max_keep.
// All comparisons that decide membership (score > thresh, iou > thresh) use round-to-nearest fp32 ops in
// the reference's order with FMA contraction disabled, so keep indices and counts are bit-exact.
#include <cooperative_groups.h>

#include "common.cuh"
#include "sort_scan.cuh"

namespace b200 {

constexpr float kScaleClamp = 4.135166556742356f;  // log(1000/16), detectron2 _DEFAULT_SCALE_CLAMP

// softmax statistics of one row (sequential over the K+1 columns: a thread owns a row)
__device__ __forceinline__ void row_stats(const float* __restrict__ row, int ncol, float* mx_out, float* sum_out) {
  float mx = -INFINITY;
  for (int k = 0; k < ncol; ++k) mx = fmaxf(mx, __ldg(row + k));
  float s = 0.f;
  for (int k = 0; k < ncol; ++k) s += expf(__ldg(row + k) - mx);
  *mx_out = mx; *sum_out = s;
}
__device__ __forceinline__ float row_prob(const float* __restrict__ row, int k, int input_is_prob, float mx, float sum) {
  const float x = __ldg(row + k);
  return input_is_prob ? x : __fdiv_rn(expf(x - mx), sum);
}

__device__ __forceinline__ float4 decode_clip(const float* __restrict__ d, float4 pb, float wx, float wy, float ww,
                                              float wh, float img_h, float img_w) {
  // Box2BoxTransform.apply_deltas, op for op (separate roundings), then Boxes.clip
  const float widths = __fsub_rn(pb.z, pb.x), heights = __fsub_rn(pb.w, pb.y);
  const float ctr_x = __fadd_rn(pb.x, __fmul_rn(0.5f, widths)), ctr_y = __fadd_rn(pb.y, __fmul_rn(0.5f, heights));
  const float dx = __fdiv_rn(d[0], wx), dy = __fdiv_rn(d[1], wy);
  const float dw = fminf(__fdiv_rn(d[2], ww), kScaleClamp), dh = fminf(__fdiv_rn(d[3], wh), kScaleClamp);
  const float pcx = __fadd_rn(__fmul_rn(dx, widths), ctr_x), pcy = __fadd_rn(__fmul_rn(dy, heights), ctr_y);
  const float pw = __fmul_rn(expf(dw), widths), ph = __fmul_rn(expf(dh), heights);
  float4 o;
  o.x = __fsub_rn(pcx, __fmul_rn(0.5f, pw)); o.y = __fsub_rn(pcy, __fmul_rn(0.5f, ph));
  o.z = __fadd_rn(pcx, __fmul_rn(0.5f, pw)); o.w = __fadd_rn(pcy, __fmul_rn(0.5f, ph));
  o.x = fminf(fmaxf(o.x, 0.f), img_w); o.y = fminf(fmaxf(o.y, 0.f), img_h);
  o.z = fminf(fmaxf(o.z, 0.f), img_w); o.w = fminf(fmaxf(o.w, 0.f), img_h);
  return o;
}

constexpr int kRowTile = 1024;  // rows handled per block iteration: one row per thread

// One CTA per image, one THREAD per ROI row: the per-row work (K+1 <= a few hundred columns) is tiny, so rows are
// the parallel axis; a block scan of the per-row candidate counts gives the torch.nonzero() order.
__global__ void __launch_bounds__(1024)
softmax_decode_compact_kernel(const float* __restrict__ scores_in, int input_is_prob, const float* __restrict__ deltas,
                              const float* __restrict__ proposals, const int32_t* __restrict__ roi_offsets,
                              const float* __restrict__ image_hw, int K, int cls_agnostic, float wx, float wy,
                              float ww, float wh, float thresh, float* __restrict__ probs_out,
                              float* __restrict__ cand_boxes, float* __restrict__ cand_scores,
                              int32_t* __restrict__ cand_roi, int32_t* __restrict__ cand_cls,
                              int32_t* __restrict__ cand_count) {
  __shared__ int s_warp[33];
  const int img = blockIdx.x;
  const int r0 = roi_offsets[img], r1 = roi_offsets[img + 1];
  const float img_h = image_hw[2 * img], img_w = image_hw[2 * img + 1];
  const int ncol = K + 1;
  const size_t out0 = (size_t)r0 * K;  // this image's candidate segment
  int running = 0;

  for (int t0 = r0; t0 < r1; t0 += kRowTile) {
    const int r = t0 + threadIdx.x;
    const bool live = r < r1;
    const float* row = scores_in + (size_t)r * ncol;
    float mx = 0.f, sum = 1.f;
    int cnt = 0;
    if (live) {
      if (!input_is_prob) row_stats(row, ncol, &mx, &sum);
      for (int k = 0; k < ncol; ++k) {
        const float p = row_prob(row, k, input_is_prob, mx, sum);
        if (probs_out) probs_out[(size_t)r * ncol + k] = p;
        cnt += (k < K && p > thresh);
      }
    }
    int total;
    const int ex = block_exclusive_scan_1024(cnt, s_warp, &total);
    if (live && cnt) {
      size_t o = out0 + running + ex;
      const float4 pb = *reinterpret_cast<const float4*>(proposals + 4 * (size_t)r);
      for (int k = 0; k < K; ++k) {
        const float p = row_prob(row, k, input_is_prob, mx, sum);   // same ops, same bits as the counting pass
        if (p > thresh) {
          const float* d = deltas + (cls_agnostic ? (size_t)r * 4 : ((size_t)r * K + k) * 4);
          *reinterpret_cast<float4*>(cand_boxes + 4 * o) = decode_clip(d, pb, wx, wy, ww, wh, img_h, img_w);
          cand_scores[o] = p;
          cand_roi[o] = r - r0;
          cand_cls[o] = k;
          ++o;
        }
      }
    }
    running += total;
  }
  if (threadIdx.x == 0) cand_count[img] = running;
}

// Many-CTA form (large proposal counts: BASELINE configs[4] sweeps up to 8192 proposals per image): 256 rows per CTA, two
// launches.  COUNT: every row's number of candidates -> per-CTA totals (and the probabilities, if asked for).  WRITE: a
// CTA's base offset is the sum of the totals of the CTAs before it in the image (<= a few dozen), the rows' offsets an
// exclusive scan inside the CTA; the candidates are recomputed with the same operations (same bits) and written in the
// torch.nonzero() order.  No host synchronisation, no atomics.
constexpr int kMcRows = 256;

__device__ __forceinline__ int block_exclusive_scan_256(int v, int* s_warp /*[9]*/, int* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int w = 0; w < kMcRows / 32; ++w) { const int t = s_warp[w]; s_warp[w] = run; run += t; }
    s_warp[8] = run;
  }
  __syncthreads();
  const int res = s_warp[warp] + inc - v;
  *total = s_warp[8];
  __syncthreads();
  return res;
}

template <bool WRITE>
__global__ void __launch_bounds__(kMcRows)
softmax_decode_compact_mc_kernel(const float* __restrict__ scores_in, int input_is_prob, const float* __restrict__ deltas,
                                 const float* __restrict__ proposals, const int32_t* __restrict__ roi_offsets,
                                 const float* __restrict__ image_hw, int K, int cls_agnostic, float wx, float wy,
                                 float ww, float wh, float thresh, float* __restrict__ probs_out,
                                 float* __restrict__ cand_boxes, float* __restrict__ cand_scores,
                                 int32_t* __restrict__ cand_roi, int32_t* __restrict__ cand_cls,
                                 int32_t* __restrict__ cand_count, int32_t* __restrict__ block_tot, int nblk) {
  __shared__ int s_warp[9];
  __shared__ int s_base;
  const int img = blockIdx.y, blk = blockIdx.x;
  const int r0 = roi_offsets[img], r1 = roi_offsets[img + 1];
  const int r = r0 + blk * kMcRows + (int)threadIdx.x;
  const bool live = r < r1;
  const int ncol = K + 1;
  const float* row = scores_in + (size_t)r * ncol;
  float mx = 0.f, sum = 1.f;
  int cnt = 0;
  if (live) {
    if (!input_is_prob) row_stats(row, ncol, &mx, &sum);
    for (int k = 0; k < ncol; ++k) {
      const float p = row_prob(row, k, input_is_prob, mx, sum);
      if (!WRITE && probs_out) probs_out[(size_t)r * ncol + k] = p;
      cnt += (k < K && p > thresh);
    }
  }
  int total;
  const int ex = block_exclusive_scan_256(cnt, s_warp, &total);
  if (!WRITE) {
    if (threadIdx.x == 0) block_tot[(size_t)img * nblk + blk] = total;
    return;
  }
  if (threadIdx.x == 0) {
    int base = 0;
    for (int b = 0; b < blk; ++b) base += block_tot[(size_t)img * nblk + b];
    s_base = base;
    if (blk == nblk - 1) cand_count[img] = base + total;
  }
  __syncthreads();
  if (live && cnt) {
    const float img_h = image_hw[2 * img], img_w = image_hw[2 * img + 1];
    size_t o = (size_t)r0 * K + s_base + ex;
    const float4 pb = *reinterpret_cast<const float4*>(proposals + 4 * (size_t)r);
    for (int k = 0; k < K; ++k) {
      const float p = row_prob(row, k, input_is_prob, mx, sum);   // same ops, same bits as the counting pass
      if (p > thresh) {
        const float* d = deltas + (cls_agnostic ? (size_t)r * 4 : ((size_t)r * K + k) * 4);
        *reinterpret_cast<float4*>(cand_boxes + 4 * o) = decode_clip(d, pb, wx, wy, ww, wh, img_h, img_w);
        cand_scores[o] = p;
        cand_roi[o] = r - r0;
        cand_cls[o] = k;
        ++o;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// NMS
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool iou_gt(const float4 a, const float4 b, float thr) {
  // torchvision nms kernel: inter / (areaA + areaB - inter) > thr, widths clamped at 0
  const float left = fmaxf(a.x, b.x), right = fminf(a.z, b.z);
  const float top = fmaxf(a.y, b.y), bottom = fminf(a.w, b.w);
  const float w = fmaxf(__fsub_rn(right, left), 0.f), h = fmaxf(__fsub_rn(bottom, top), 0.f);
  // disjoint boxes (the vast majority of pairs): inter == 0 exactly, and 0/union (or 0/0 = NaN) is never > thr for
  // thr >= 0 — skip the IEEE division without changing any decision
  if (thr >= 0.f && (w == 0.f || h == 0.f)) return false;
  const float inter = __fmul_rn(w, h);
  const float sa = __fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y));
  const float sb = __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
  return __fdiv_rn(inter, __fsub_rn(__fadd_rn(sa, sb), inter)) > thr;
}

struct NmsWorkspace {
  float* max1;                 // (N)   boxes.max()+1 per segment (0 when the >=40000 path applies)
  int32_t* class_start;        // (N, num_classes+1)
  int32_t* order;              // (total_capacity) candidate indices grouped by class (segment-relative)
  unsigned long long* kept;    // (total_capacity) survivor keys, ~0 = suppressed
  unsigned long long* scratch; // (total_capacity rounded up to pow2 per use) global sort fallback
};

constexpr int kPrepThreads = 1024;
constexpr int kMaxClasses = 1024;

__global__ void __launch_bounds__(kPrepThreads)
nms_prepare_kernel(const float* __restrict__ boxes, const int32_t* __restrict__ classes,
                   const int32_t* __restrict__ seg_offsets, const int32_t* __restrict__ seg_count, int num_classes,
                   float* __restrict__ max1, int32_t* __restrict__ class_start, int32_t* __restrict__ order) {
  extern __shared__ int s_dyn[];          // hist[num_classes] | cursor[num_classes] | wcnt[32][num_classes]
  int* s_hist = s_dyn;
  int* s_cursor = s_dyn + num_classes;
  int* s_wcnt = s_dyn + 2 * num_classes;
  __shared__ float s_red[32];
  const int img = blockIdx.x;
  const int base = seg_offsets[img], n = seg_count[img];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* bx = boxes + 4 * (size_t)base;
  const int32_t* cl = classes + base;

  for (int c = threadIdx.x; c < num_classes; c += blockDim.x) s_hist[c] = 0;
  __syncthreads();
  float mx = -INFINITY;
  for (int i = threadIdx.x; i < 4 * n; i += blockDim.x) mx = fmaxf(mx, bx[i]);
  for (int i = threadIdx.x; i < n; i += blockDim.x) atomicAdd(&s_hist[cl[i]], 1);
  mx = warp_max(mx);
  if (lane == 0) s_red[warp] = mx;
  __syncthreads();
  if (warp == 0) {
    mx = warp_max(s_red[lane]);
    // detectron2 0.3 batched_nms: coordinate trick only below 40000 boxes
    if (lane == 0) max1[img] = (n > 0 && n < 40000) ? __fadd_rn(mx, 1.0f) : 0.f;
  }
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int c = 0; c < num_classes; ++c) {
      class_start[(size_t)img * (num_classes + 1) + c] = acc;
      s_cursor[c] = acc;
      acc += s_hist[c];
    }
    class_start[(size_t)img * (num_classes + 1) + num_classes] = acc;
  }
  __syncthreads();
  // stable placement, 1024 candidates per round
  for (int i0 = 0; i0 < n; i0 += blockDim.x) {
    for (int k = threadIdx.x; k < 32 * num_classes; k += blockDim.x) s_wcnt[k] = 0;
    __syncthreads();
    const int i = i0 + threadIdx.x;
    const bool valid = i < n;
    const int c = valid ? cl[i] : -1;
    const unsigned peers = __match_any_sync(0xffffffffu, c);
    const int rank = __popc(peers & ((1u << lane) - 1u));
    if (valid && rank == 0) s_wcnt[warp * num_classes + c] = __popc(peers);
    __syncthreads();
    for (int cc = threadIdx.x; cc < num_classes; cc += blockDim.x) {
      int run = s_cursor[cc];
      for (int w = 0; w < 32; ++w) {
        const int t = s_wcnt[w * num_classes + cc];
        s_wcnt[w * num_classes + cc] = run;
        run += t;
      }
      s_cursor[cc] = run;
    }
    __syncthreads();
    if (valid) order[base + s_wcnt[warp * num_classes + c] + rank] = i;
    __syncthreads();
  }
}

constexpr int kNmsThreads = 256;
constexpr int kNmsSmemBoxes = 4096;  // per (class, image) handled fully in shared memory by the 256-thread kernel
// Larger slices (one class taking most of an image's 4096-8192 proposals: the BASELINE configs[4] sweep with an untrained
// classifier) go to a second instantiation: 1024 threads, 8192 boxes x 24 B + bitmap = 197 KB of shared memory, so the
// bitonic sort and the sweep stay on chip; beyond that the global-memory path of the same kernel.
constexpr int kNmsBigBoxes = 8192;
constexpr int kNmsBigThreads = 1024;
// PRESORTED variant (RPN proposal selection, rpn_select.cu): the candidates of a class already arrive in
// (score desc, index asc) order, so no keys are built or sorted and shared memory holds boxes + bitmap only:
// 12288 boxes x 16 B + 1.5 KB = 198 KB, one 1024-thread CTA per (level, image).
constexpr int kNmsPresortedBoxes = 12288;
constexpr int kNmsPresortedThreads = 1024;
constexpr int kNmsCluster = 8;       // portable cluster size

template <int THREADS, bool PRESORTED, int CAP>
__global__ void __launch_bounds__(THREADS)
nms_class_kernel(const float* __restrict__ boxes, const float* __restrict__ scores,
                 const int32_t* __restrict__ seg_offsets, int num_classes, float thr,
                 const float* __restrict__ max1, const int32_t* __restrict__ class_start,
                 const int32_t* __restrict__ order, unsigned long long* __restrict__ kept,
                 unsigned long long* __restrict__ scratch, int max_keep, int skip_upto, int skip_above) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  __shared__ unsigned long long s_diag[64];
  __shared__ unsigned long long s_keep64;
  const int c = blockIdx.x, img = blockIdx.y;
  const int base = seg_offsets[img];
  const int cs = class_start[(size_t)img * (num_classes + 1) + c];
  const int m = class_start[(size_t)img * (num_classes + 1) + c + 1] - cs;
  // slices up to skip_upto / above skip_above (> 0) belong to another launch of this call
  if (m == 0 || m <= skip_upto || (skip_above > 0 && m > skip_above)) return;
  const float off = __fmul_rn((float)c, max1[img]);
  const int32_t* ord = order + base + cs;
  unsigned long long* kept_out = kept + base + cs;

  if (m == 1) {
    if (threadIdx.x == 0) kept_out[0] = ((unsigned long long)desc_key(scores[base + ord[0]]) << 32) | (uint32_t)ord[0];
    return;
  }
  const int m2 = next_pow2(m);
  constexpr int kCap = CAP;
  const bool in_smem = m <= kCap;
  unsigned long long* gslice = scratch + 4 * (size_t)(base + cs);  // 4*m u64 per class slice (host sizes scratch 4x)
  unsigned long long* keys = PRESORTED ? nullptr : (in_smem ? reinterpret_cast<unsigned long long*>(s_raw) : gslice);
  float4* sbox = reinterpret_cast<float4*>(s_raw + (PRESORTED ? 0 : (size_t)kCap * 8));
  uint32_t* removed = reinterpret_cast<uint32_t*>(s_raw + (size_t)kCap * (PRESORTED ? 16 : 24));

  if (!PRESORTED) {
    // (score desc, position-in-class asc); positions are ascending candidate index (stable counting sort)
    for (int p = threadIdx.x; p < m2; p += blockDim.x)
      keys[p] = p < m ? (((unsigned long long)desc_key(scores[base + ord[p]]) << 32) | (uint32_t)p) : ~0ull;
    __syncthreads();
    bitonic_sort_u64(keys, m2);
  }
  // position in the class slice of the j-th box in (score desc, index asc) order
  auto pos_of = [&](int j) -> int { return PRESORTED ? j : (int)(keys[j] & 0xffffffffu); };

  auto load_box = [&](int j) -> float4 {
    const int cand = ord[pos_of(j)];
    float4 b = *reinterpret_cast<const float4*>(boxes + 4 * (size_t)(base + cand));
    b.x = __fadd_rn(b.x, off); b.y = __fadd_rn(b.y, off); b.z = __fadd_rn(b.z, off); b.w = __fadd_rn(b.w, off);
    return b;
  };
  if (in_smem) {
    for (int j = threadIdx.x; j < m; j += blockDim.x) sbox[j] = load_box(j);
    for (int j = threadIdx.x; j < (m + 31) / 32; j += blockDim.x) removed[j] = 0u;
  } else {
    // global path: `removed` bitmap lives in the upper half of this class's scratch slice
    removed = reinterpret_cast<uint32_t*>(gslice + 2 * (size_t)m);
    for (int j = threadIdx.x; j < (m + 31) / 32; j += blockDim.x) removed[j] = 0u;
  }
  __syncthreads();
  auto get_box = [&](int j) -> float4 { return in_smem ? sbox[j] : load_box(j); };

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int nkept = 0;
  for (int b0 = 0; b0 < m; b0 += 64) {
    const int nb = min(64, m - b0);
    // diagonal 64x64 block, all 8 warps: warp -> row t, lane -> columns (lane, lane+32); ballots assemble the
    // 64-bit row mask "boxes later in the block that box t suppresses"
    for (int t = warp; t < 64; t += THREADS / 32) {
      unsigned m0 = 0u, m1 = 0u;
      if (t < nb) {
        const float4 a = get_box(b0 + t);
        const int j0 = lane, j1 = lane + 32;
        const bool h0 = j0 > t && j0 < nb && iou_gt(a, get_box(b0 + j0), thr);
        const bool h1 = j1 > t && j1 < nb && iou_gt(a, get_box(b0 + j1), thr);
        m0 = __ballot_sync(0xffffffffu, h0);
        m1 = __ballot_sync(0xffffffffu, h1);
      }
      if (lane == 0) s_diag[t] = (unsigned long long)m0 | ((unsigned long long)m1 << 32);
    }
    __syncthreads();
    if (warp == 0) {
      // serial dependency resolved from registers: lane holds rows (lane, lane+32); the row in turn is broadcast
      const unsigned long long lo = s_diag[lane], hi = s_diag[lane + 32];
      unsigned long long dead = (unsigned long long)removed[b0 >> 5] |
                                ((b0 + 32 < m) ? ((unsigned long long)removed[(b0 >> 5) + 1] << 32) : 0ull);
      unsigned long long keep = 0ull;
      for (int t = 0; t < nb; ++t) {
        const unsigned long long row = __shfl_sync(0xffffffffu, t < 32 ? lo : hi, t & 31);
        if (!((dead >> t) & 1ull)) { keep |= 1ull << t; dead |= row; }
      }
      if (lane == 0) s_keep64 = keep;
    }
    __syncthreads();
    unsigned long long keep = s_keep64;
    // only the image's first max_keep survivors can reach the output: a class never needs more than that
    const int room = max_keep - nkept;
    if (__popcll(keep) > room) {
      int seen = 0;
      unsigned long long trimmed = 0ull;
      for (int t = 0; t < nb && seen < room; ++t)
        if ((keep >> t) & 1ull) { trimmed |= 1ull << t; ++seen; }
      keep = trimmed;
    }
    nkept += __popcll(keep);
    if ((int)threadIdx.x < nb) {
      const int j = b0 + threadIdx.x;
      const uint32_t cand = (uint32_t)ord[pos_of(j)];
      const unsigned long long hi = PRESORTED ? ((unsigned long long)desc_key(scores[base + cand]) << 32)
                                              : (keys[j] & 0xffffffff00000000ull);
      kept_out[j] = ((keep >> threadIdx.x) & 1ull) ? (hi | cand) : ~0ull;
    }
    if (nkept >= max_keep) {
      for (int j = b0 + 64 + threadIdx.x; j < m; j += blockDim.x) kept_out[j] = ~0ull;
      break;                                            // uniform: every thread computed the same nkept
    }
    if (keep != 0ull) {
      for (int j = b0 + 64 + threadIdx.x; j < m; j += blockDim.x) {
        if ((removed[j >> 5] >> (j & 31)) & 1u) continue;
        const float4 bj = get_box(j);
        unsigned long long kk = keep;
        bool dead = false;
        while (kk && !dead) {
          const int t = __ffsll((long long)kk) - 1;
          kk &= kk - 1;
          dead = iou_gt(get_box(b0 + t), bj, thr);
        }
        if (dead) atomicOr(&removed[j >> 5], 1u << (j & 31));
      }
    }
    __syncthreads();
  }
}

// PRESORTED slices of up to kNmsPresortedBoxes boxes on a thread-block CLUSTER (RPN proposal selection: one level of one
// image is 6000-12000 boxes, far more survivor x box work than one SM issues in reasonable time).  Every CTA of the
// cluster keeps all boxes of the slice in its own shared memory and resolves each 64-box diagonal block redundantly (same
// inputs, same result), but sweeps the survivors only over the bitmap words it owns (word w belongs to rank w % CL).
// The two `removed` words a diagonal block needs are read from their owners through distributed shared memory; one
// cluster barrier per block orders the sweeps before those reads.  Rank 0 writes the result.
template <int CL>
__global__ void __launch_bounds__(kNmsPresortedThreads)
nms_presorted_cluster_kernel(const float* __restrict__ boxes, const float* __restrict__ scores,
                             const int32_t* __restrict__ seg_offsets, int num_classes, float thr,
                             const float* __restrict__ max1, const int32_t* __restrict__ class_start,
                             const int32_t* __restrict__ order, unsigned long long* __restrict__ kept, int max_keep) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ __align__(16) unsigned char s_raw[];
  __shared__ unsigned long long s_diag[64];
  __shared__ unsigned long long s_keep64;
  __shared__ uint32_t s_dead[2];
  const int rank = (int)cluster.block_rank();
  const int c = blockIdx.x / CL, img = blockIdx.y;
  const int base = seg_offsets[img];
  const int cs = class_start[(size_t)img * (num_classes + 1) + c];
  const int m = class_start[(size_t)img * (num_classes + 1) + c + 1] - cs;
  if (m == 0 || m > kNmsPresortedBoxes) return;              // uniform over the cluster; larger slices: fallback kernel
  const float off = __fmul_rn((float)c, max1[img]);
  const int32_t* ord = order + base + cs;
  unsigned long long* kept_out = kept + base + cs;
  float4* sbox = reinterpret_cast<float4*>(s_raw);
  uint32_t* removed = reinterpret_cast<uint32_t*>(s_raw + (size_t)kNmsPresortedBoxes * 16);

  for (int j = threadIdx.x; j < m; j += blockDim.x) {
    float4 b = *reinterpret_cast<const float4*>(boxes + 4 * (size_t)(base + ord[j]));
    b.x = __fadd_rn(b.x, off); b.y = __fadd_rn(b.y, off); b.z = __fadd_rn(b.z, off); b.w = __fadd_rn(b.w, off);
    sbox[j] = b;
  }
  for (int j = threadIdx.x; j < kNmsPresortedBoxes / 32; j += blockDim.x) removed[j] = 0u;
  cluster.sync();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int nkept = 0;
  for (int b0 = 0; b0 < m; b0 += 64) {
    const int nb = min(64, m - b0);
    if (threadIdx.x < 2) {
      const int w = (b0 >> 5) + threadIdx.x;
      s_dead[threadIdx.x] = (w * 32 < m) ? *cluster.map_shared_rank(removed + w, w % CL) : 0u;
    }
    for (int t = warp; t < 64; t += kNmsPresortedThreads / 32) {
      unsigned m0 = 0u, m1 = 0u;
      if (t < nb) {
        const float4 a = sbox[b0 + t];
        const int j0 = lane, j1 = lane + 32;
        const bool h0 = j0 > t && j0 < nb && iou_gt(a, sbox[b0 + j0], thr);
        const bool h1 = j1 > t && j1 < nb && iou_gt(a, sbox[b0 + j1], thr);
        m0 = __ballot_sync(0xffffffffu, h0);
        m1 = __ballot_sync(0xffffffffu, h1);
      }
      if (lane == 0) s_diag[t] = (unsigned long long)m0 | ((unsigned long long)m1 << 32);
    }
    __syncthreads();
    if (warp == 0) {
      const unsigned long long lo = s_diag[lane], hi = s_diag[lane + 32];
      unsigned long long dead = (unsigned long long)s_dead[0] | ((unsigned long long)s_dead[1] << 32);
      unsigned long long keep = 0ull;
      for (int t = 0; t < nb; ++t) {
        const unsigned long long row = __shfl_sync(0xffffffffu, t < 32 ? lo : hi, t & 31);
        if (!((dead >> t) & 1ull)) { keep |= 1ull << t; dead |= row; }
      }
      if (lane == 0) s_keep64 = keep;
    }
    __syncthreads();
    unsigned long long keep = s_keep64;
    const int room = max_keep - nkept;
    if (__popcll(keep) > room) {
      int seen = 0;
      unsigned long long trimmed = 0ull;
      for (int t = 0; t < nb && seen < room; ++t)
        if ((keep >> t) & 1ull) { trimmed |= 1ull << t; ++seen; }
      keep = trimmed;
    }
    nkept += __popcll(keep);
    if (rank == 0 && (int)threadIdx.x < nb) {
      const int j = b0 + threadIdx.x;
      const uint32_t cand = (uint32_t)ord[j];
      kept_out[j] = ((keep >> threadIdx.x) & 1ull) ? (((unsigned long long)desc_key(scores[base + cand]) << 32) | cand) : ~0ull;
    }
    if (nkept >= max_keep) {                                   // identical in every CTA of the cluster
      if (rank == 0)
        for (int j = b0 + 64 + threadIdx.x; j < m; j += blockDim.x) kept_out[j] = ~0ull;
      break;
    }
    if (keep != 0ull) {
      // own words: w = w_first + CL * i, one box per thread and step
      const int wlo = (b0 + 64) >> 5;
      const int w_first = wlo + ((rank - wlo) % CL + CL) % CL;
      for (int idx = threadIdx.x;; idx += blockDim.x) {
        const int w = w_first + CL * (idx >> 5);
        const int j = w * 32 + (idx & 31);
        if (w * 32 >= m) break;
        if (j >= m || ((removed[w] >> (j & 31)) & 1u)) continue;
        const float4 bj = sbox[j];
        unsigned long long kk = keep;
        bool dead = false;
        while (kk && !dead) {
          const int t = __ffsll((long long)kk) - 1;
          kk &= kk - 1;
          dead = iou_gt(sbox[b0 + t], bj, thr);
        }
        if (dead) atomicOr(&removed[w], 1u << (j & 31));
      }
    }
    cluster.sync();
  }
  cluster.sync();                                              // nobody leaves while a peer may still read its bitmap
}

constexpr int kFinalSmemKeys = 4096;

__global__ void __launch_bounds__(1024)
nms_finalize_kernel(const int32_t* __restrict__ seg_offsets, const int32_t* __restrict__ seg_count,
                    unsigned long long* __restrict__ kept, unsigned long long* __restrict__ scratch, int max_keep,
                    int32_t* __restrict__ keep, int32_t* __restrict__ keep_count) {
  __shared__ unsigned long long s_keys[kFinalSmemKeys];
  __shared__ int s_n;
  const int img = blockIdx.x;
  const int base = seg_offsets[img], n = seg_count[img];
  if (threadIdx.x == 0) s_n = 0;
  __syncthreads();
  // count survivors
  int local = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) local += kept[base + i] != ~0ull;
  local = (int)warp_sum((float)local);  // n < 2^24 so exact in fp32
  if ((threadIdx.x & 31) == 0 && local) atomicAdd(&s_n, local);
  __syncthreads();
  const int nk = s_n;
  __syncthreads();
  unsigned long long* keys = nk <= kFinalSmemKeys ? s_keys : scratch + 4 * (size_t)base;
  const int n2 = next_pow2(max(nk, 1));
  if (threadIdx.x == 0) s_n = 0;
  __syncthreads();
  // compaction order is irrelevant: keys are unique and get sorted next
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const unsigned long long k = kept[base + i];
    if (k != ~0ull) keys[atomicAdd(&s_n, 1)] = k;
  }
  for (int i = nk + threadIdx.x; i < n2; i += blockDim.x) keys[i] = ~0ull;
  __syncthreads();
  bitonic_sort_u64(keys, n2);
  const int out_n = max_keep >= 0 ? min(nk, max_keep) : nk;
  for (int i = threadIdx.x; i < out_n; i += blockDim.x) keep[(size_t)img * max(max_keep, 0) + i] = (int32_t)(keys[i] & 0xffffffffu);
  if (threadIdx.x == 0) keep_count[img] = out_n;
}

__global__ void gather_detections_kernel(const float* __restrict__ cand_boxes, const float* __restrict__ cand_scores,
                                         const int32_t* __restrict__ cand_roi, const int32_t* __restrict__ cand_cls,
                                         const int32_t* __restrict__ seg_offsets, const int32_t* __restrict__ keep,
                                         const int32_t* __restrict__ keep_count, int max_keep,
                                         float* __restrict__ out_boxes, float* __restrict__ out_scores,
                                         int64_t* __restrict__ out_classes, int64_t* __restrict__ out_roi) {
  const int img = blockIdx.x;
  const int nk = keep_count[img], base = seg_offsets[img];
  for (int i = threadIdx.x; i < max_keep; i += blockDim.x) {
    const size_t o = (size_t)img * max_keep + i;
    if (i < nk) {
      const int j = base + keep[o];
      *reinterpret_cast<float4*>(out_boxes + 4 * o) = *reinterpret_cast<const float4*>(cand_boxes + 4 * (size_t)j);
      out_scores[o] = cand_scores[j];
      out_classes[o] = cand_cls[j];
      out_roi[o] = cand_roi[j];
    } else {
      *reinterpret_cast<float4*>(out_boxes + 4 * o) = make_float4(0.f, 0.f, 0.f, 0.f);
      out_scores[o] = 0.f;
      out_classes[o] = -1;
      out_roi[o] = -1;
    }
  }
}

// detectron2 0.3 detector_postprocess (called from defrcn/modeling/meta_arch/rcnn.py:69-73) on the padded
// detection tensors: scale to the requested output resolution, clip, drop empty boxes (ordered, in place).
__global__ void __launch_bounds__(1024)
detector_postprocess_kernel(float* __restrict__ boxes, float* __restrict__ scores, int64_t* __restrict__ classes,
                            int64_t* __restrict__ roi_inds, int32_t* __restrict__ counts,
                            const float* __restrict__ scale_xy, const float* __restrict__ out_hw, int max_keep) {
  __shared__ int s_warp[33];
  const int img = blockIdx.x;
  const int n = min(counts[img], max_keep);
  const float sx = scale_xy[2 * img], sy = scale_xy[2 * img + 1];
  const float oh = out_hw[2 * img], ow = out_hw[2 * img + 1];
  const size_t base = (size_t)img * max_keep;
  int running = 0;
  for (int t0 = 0; t0 < n; t0 += 1024) {
    const int i = t0 + threadIdx.x;
    float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
    float sc = 0.f;
    int64_t c = -1, r = -1;
    int keepf = 0;
    if (i < n) {
      b = *reinterpret_cast<const float4*>(boxes + 4 * (base + i));
      sc = scores[base + i];
      if (classes) c = classes[base + i];
      if (roi_inds) r = roi_inds[base + i];
      b.x = fminf(fmaxf(__fmul_rn(b.x, sx), 0.f), ow); b.y = fminf(fmaxf(__fmul_rn(b.y, sy), 0.f), oh);
      b.z = fminf(fmaxf(__fmul_rn(b.z, sx), 0.f), ow); b.w = fminf(fmaxf(__fmul_rn(b.w, sy), 0.f), oh);
      keepf = (__fsub_rn(b.z, b.x) > 0.f) && (__fsub_rn(b.w, b.y) > 0.f);
    }
    int total;
    const int ex = block_exclusive_scan_1024(keepf, s_warp, &total);   // every read of this tile precedes its writes
    if (keepf) {
      const size_t o = base + running + ex;
      *reinterpret_cast<float4*>(boxes + 4 * o) = b;
      scores[o] = sc;
      if (classes) classes[o] = c;
      if (roi_inds) roi_inds[o] = r;
    }
    running += total;
  }
  __syncthreads();
  for (int i = running + threadIdx.x; i < n; i += 1024) {
    *reinterpret_cast<float4*>(boxes + 4 * (base + i)) = make_float4(0.f, 0.f, 0.f, 0.f);
    scores[base + i] = 0.f;
    if (classes) classes[base + i] = -1;
    if (roi_inds) roi_inds[base + i] = -1;
  }
  if (threadIdx.x == 0) counts[img] = running;
}

static NmsWorkspace carve(void* ws, int N, int total_capacity, int num_classes) {
  NmsWorkspace w;
  unsigned char* p = (unsigned char*)ws;
  w.kept = (unsigned long long*)p;    p += align_up((size_t)total_capacity * 8, 256);
  w.scratch = (unsigned long long*)p; p += align_up((size_t)total_capacity * 8 * 4, 256);
  w.order = (int32_t*)p;              p += align_up((size_t)total_capacity * 4, 256);
  w.class_start = (int32_t*)p;        p += align_up((size_t)N * (num_classes + 1) * 4, 256);
  w.max1 = (float*)p;
  return w;
}

}  // namespace b200

using namespace b200;

extern "C" size_t b200_softmax_decode_compact_workspace_bytes(int N, int max_rois_per_image) {
  if (N <= 0 || max_rois_per_image <= 0) return 0;
  return (size_t)N * ceil_div(max_rois_per_image, kMcRows) * sizeof(int32_t);
}

extern "C" int b200_softmax_decode_compact(const float* scores_in, int input_is_prob, const float* deltas,
                                           const float* proposals, const int32_t* roi_offsets, const float* image_hw,
                                           int N, int R, int K, int cls_agnostic, float wx, float wy, float ww,
                                           float wh, float score_thresh, float* probs_out, float* cand_boxes,
                                           float* cand_scores, int32_t* cand_roi, int32_t* cand_cls,
                                           int32_t* cand_count, int max_rois_per_image, void* workspace,
                                           size_t workspace_bytes, b200_stream_t stream) {
  B200_CHECK_ARG(N >= 0 && R >= 0 && K > 0, "softmax_decode_compact: bad shape");
  B200_CHECK_ARG(roi_offsets && image_hw && cand_count, "softmax_decode_compact: null index tensors");
  B200_CHECK_ARG(R == 0 || (scores_in && deltas && proposals && cand_boxes && cand_scores && cand_roi && cand_cls),
                 "softmax_decode_compact: null tensor");
  if (N == 0) return B200_OK;
  if (workspace && max_rois_per_image > 0) {
    // many CTAs per image (count, then write); max_rois_per_image bounds roi_offsets[i + 1] - roi_offsets[i]
    const int nblk = ceil_div(max_rois_per_image, kMcRows);
    if (workspace_bytes < b200_softmax_decode_compact_workspace_bytes(N, max_rois_per_image)) {
      set_error("softmax_decode_compact: workspace too small");
      return B200_ERR_WORKSPACE;
    }
    dim3 grid(nblk, N);
    softmax_decode_compact_mc_kernel<false><<<grid, kMcRows, 0, (cudaStream_t)stream>>>(
        scores_in, input_is_prob, deltas, proposals, roi_offsets, image_hw, K, cls_agnostic, wx, wy, ww, wh, score_thresh,
        probs_out, cand_boxes, cand_scores, cand_roi, cand_cls, cand_count, (int32_t*)workspace, nblk);
    softmax_decode_compact_mc_kernel<true><<<grid, kMcRows, 0, (cudaStream_t)stream>>>(
        scores_in, input_is_prob, deltas, proposals, roi_offsets, image_hw, K, cls_agnostic, wx, wy, ww, wh, score_thresh,
        probs_out, cand_boxes, cand_scores, cand_roi, cand_cls, cand_count, (int32_t*)workspace, nblk);
    B200_CUDA_LAUNCH_CHECK("softmax_decode_compact (many-CTA)");
    return B200_OK;
  }
  softmax_decode_compact_kernel<<<N, 1024, 0, (cudaStream_t)stream>>>(
      scores_in, input_is_prob, deltas, proposals, roi_offsets, image_hw, K, cls_agnostic, wx, wy, ww, wh,
      score_thresh, probs_out, cand_boxes, cand_scores, cand_roi, cand_cls, cand_count);
  B200_CUDA_LAUNCH_CHECK("softmax_decode_compact");
  return B200_OK;
}

extern "C" size_t b200_batched_nms_workspace_bytes(int N, int total_capacity, int num_classes) {
  return align_up((size_t)total_capacity * 8, 256) + align_up((size_t)total_capacity * 32, 256) +
         align_up((size_t)total_capacity * 4, 256) + align_up((size_t)N * (num_classes + 1) * 4, 256) +
         align_up((size_t)N * 4, 256);
}

namespace b200 {
// shared by b200_batched_nms and b200_rpn_select_proposals (rpn_select.cu); `presorted` = within every class the
// candidates already are in (score desc, index asc) order; max_slice_hint = upper bound of a class slice (0 = unknown)
int run_batched_nms(const float* boxes, const float* scores, const int32_t* classes, const int32_t* seg_offsets,
                    const int32_t* seg_count, int N, int total_capacity, int num_classes, float iou_thresh,
                    int max_keep, int32_t* keep, int32_t* keep_count, void* workspace, size_t workspace_bytes,
                    bool presorted, int max_slice_hint, cudaStream_t st) {
  if (!workspace || workspace_bytes < b200_batched_nms_workspace_bytes(N, total_capacity, num_classes)) {
    set_error("batched_nms: workspace too small");
    return B200_ERR_WORKSPACE;
  }
  NmsWorkspace w = carve(workspace, N, total_capacity, num_classes);
  const size_t prep_smem = (size_t)(2 + 32) * num_classes * sizeof(int);
  if (prep_smem > 48 * 1024)
    B200_CUDA_CALL(cudaFuncSetAttribute(nms_prepare_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)prep_smem));
  nms_prepare_kernel<<<N, kPrepThreads, prep_smem, st>>>(boxes, classes, seg_offsets, seg_count, num_classes, w.max1,
                                                        w.class_start, w.order);
  B200_CUDA_LAUNCH_CHECK("nms_prepare");
  // `scratch` gives every class slice 4x its length: 2x for pow2 padding of the keys, 2x for the bitmap.
  // slices are addressed at (base+cs)*4 to keep them disjoint.
  if (presorted) {
    // shifted boxes (16 B) + removed bitmap; slices <= kNmsPresortedBoxes on a cluster of kNmsCluster CTAs, larger ones
    // (not reachable from b200_rpn_select_proposals with pre_nms_topk <= 12288) on the single-CTA kernel's global path
    const size_t cls_smem = (size_t)kNmsPresortedBoxes * 16 + kNmsPresortedBoxes / 8;
    auto kc = nms_presorted_cluster_kernel<kNmsCluster>;
    auto k = nms_class_kernel<kNmsPresortedThreads, true, kNmsPresortedBoxes>;
    // set on every call: the attribute is per device and the call is cheap (no process-wide flag to race on)
    B200_CUDA_CALL(cudaFuncSetAttribute(kc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cls_smem));
    B200_CUDA_CALL(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cls_smem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(kNmsCluster * num_classes, N);
    cfg.blockDim = dim3(kNmsPresortedThreads);
    cfg.dynamicSmemBytes = cls_smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kNmsCluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    B200_CUDA_CALL(cudaLaunchKernelEx(&cfg, kc, boxes, scores, seg_offsets, num_classes, iou_thresh, (const float*)w.max1,
                                      (const int32_t*)w.class_start, (const int32_t*)w.order, w.kept, max_keep));
    if ((max_slice_hint > 0 ? max_slice_hint : total_capacity) > kNmsPresortedBoxes)
      k<<<dim3(num_classes, N), kNmsPresortedThreads, cls_smem, st>>>(boxes, scores, seg_offsets, num_classes, iou_thresh,
                                                                     w.max1, w.class_start, w.order, w.kept, w.scratch,
                                                                     max_keep, kNmsPresortedBoxes, 0);
  } else {
    // keys (8 B) + shifted boxes (16 B) + removed bitmap.  A class slice cannot exceed max_slice_hint boxes (the image's ROI
    // count, when the caller knows it): only then-possible large slices cost the second launch.
    const int bound = max_slice_hint > 0 ? min(max_slice_hint, total_capacity) : total_capacity;
    const bool big = bound > kNmsSmemBoxes;
    const size_t cls_smem = (size_t)kNmsSmemBoxes * 24 + kNmsSmemBoxes / 8;
    auto k = nms_class_kernel<kNmsThreads, false, kNmsSmemBoxes>;
    // set on every call: the attribute is per device and the call is cheap (no process-wide flag to race on)
    B200_CUDA_CALL(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cls_smem));
    k<<<dim3(num_classes, N), kNmsThreads, cls_smem, st>>>(boxes, scores, seg_offsets, num_classes, iou_thresh, w.max1,
                                                          w.class_start, w.order, w.kept, w.scratch, max_keep, 0,
                                                          big ? kNmsSmemBoxes : 0);
    if (big) {
      B200_CUDA_LAUNCH_CHECK("nms_class");
      const size_t big_smem = (size_t)kNmsBigBoxes * 24 + kNmsBigBoxes / 8;
      auto kb = nms_class_kernel<kNmsBigThreads, false, kNmsBigBoxes>;
      B200_CUDA_CALL(cudaFuncSetAttribute(kb, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)big_smem));
      kb<<<dim3(num_classes, N), kNmsBigThreads, big_smem, st>>>(boxes, scores, seg_offsets, num_classes, iou_thresh, w.max1,
                                                                w.class_start, w.order, w.kept, w.scratch, max_keep,
                                                                kNmsSmemBoxes, 0);
    }
  }
  B200_CUDA_LAUNCH_CHECK("nms_class");
  nms_finalize_kernel<<<N, 1024, 0, st>>>(seg_offsets, seg_count, w.kept, w.scratch, max_keep, keep, keep_count);
  B200_CUDA_LAUNCH_CHECK("nms_finalize");
  return B200_OK;
}
}  // namespace b200

extern "C" int b200_batched_nms(const float* boxes, const float* scores, const int32_t* classes,
                                const int32_t* seg_offsets, const int32_t* seg_count, int N, int total_capacity,
                                int num_classes, float iou_thresh, int max_keep, int max_class_slice, int32_t* keep,
                                int32_t* keep_count, void* workspace, size_t workspace_bytes, b200_stream_t stream) {
  B200_CHECK_ARG(N >= 0 && total_capacity >= 0 && num_classes > 0 && num_classes <= kMaxClasses && max_class_slice >= 0,
                 "batched_nms: bad shape");
  B200_CHECK_ARG(max_keep >= 0, "batched_nms: max_keep must be >= 0 (pass the capacity for 'all')");
  B200_CHECK_ARG(seg_offsets && seg_count && keep_count && (keep || max_keep == 0), "batched_nms: null tensor");
  if (N == 0) return B200_OK;
  return run_batched_nms(boxes, scores, classes, seg_offsets, seg_count, N, total_capacity, num_classes, iou_thresh,
                         max_keep, keep, keep_count, workspace, workspace_bytes, false, max_class_slice, (cudaStream_t)stream);
}

extern "C" int b200_gather_detections(const float* cand_boxes, const float* cand_scores, const int32_t* cand_roi,
                                      const int32_t* cand_cls, const int32_t* seg_offsets, const int32_t* keep,
                                      const int32_t* keep_count, int N, int max_keep, float* out_boxes,
                                      float* out_scores, int64_t* out_classes, int64_t* out_roi_inds,
                                      b200_stream_t stream) {
  B200_CHECK_ARG(N >= 0 && max_keep >= 0, "gather_detections: bad shape");
  if (N == 0 || max_keep == 0) return B200_OK;
  gather_detections_kernel<<<N, 128, 0, (cudaStream_t)stream>>>(cand_boxes, cand_scores, cand_roi, cand_cls, seg_offsets,
                                                               keep, keep_count, max_keep, out_boxes, out_scores,
                                                               out_classes, out_roi_inds);
  B200_CUDA_LAUNCH_CHECK("gather_detections");
  return B200_OK;
}

extern "C" int b200_detector_postprocess(float* boxes, float* scores, int64_t* classes, int64_t* roi_inds,
                                         int32_t* counts, const float* scale_xy, const float* out_hw, int N,
                                         int max_keep, b200_stream_t stream) {
  B200_CHECK_ARG(N >= 0 && max_keep >= 0, "detector_postprocess: bad shape");
  if (N == 0 || max_keep == 0) return B200_OK;
  B200_CHECK_ARG(boxes && scores && counts && scale_xy && out_hw, "detector_postprocess: null tensor");
  detector_postprocess_kernel<<<N, 1024, 0, (cudaStream_t)stream>>>(boxes, scores, classes, roi_inds, counts, scale_xy,
                                                                   out_hw, max_keep);
  B200_CUDA_LAUNCH_CHECK("detector_postprocess");
  return B200_OK;
}
