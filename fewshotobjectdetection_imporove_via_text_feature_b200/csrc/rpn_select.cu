// RPN proposal selection on the device, no host synchronisation — the stage right before the ROI head.
//   reference: detectron2 0.3 find_top_rpn_proposals, vendored at
//   defrcn/modeling/proposal_generator/proposal_utils.py:13-118 (per level: sort objectness, keep pre_nms_topk; per
//   image: drop non-finite, clip, drop boxes not larger than min_box_size, batched_nms by level, keep post_nms_topk).
//
// Kernels
//   rpn_topk_filter_kernel : one 1024-thread CTA per image, levels in turn.  An MSB-first radix select over the
//       order-preserving integer image of the logits finds the pre_nms_topk-th key (four 8-bit histogram passes, no
//       full sort of the ~30-180 k anchors); the selected (key, anchor) pairs are sorted by a shared-memory bitonic
//       network (<= 16384 keys, 128 KB); then, in sorted order, finite check + clip + min-size test + ORDERED
//       compaction (block scan) write the image's candidate segment.  Ties keep the lower anchor index first, i.e.
//       the order of the stable descending sort the reference's CPU path performs.
//   nms_presorted_cluster_kernel (detect_post.cu) : the candidates of a level arrive sorted, so no key sort; a cluster
//       of 8 CTAs per (level, image) keeps all <= 12288 boxes in each CTA's shared memory, resolves the 64-box diagonal
//       blocks redundantly and splits the survivor sweep by bitmap word, exchanging two words per block through
//       distributed shared memory.
//   rpn_gather_kernel      : kept candidates -> padded (N, post_nms_topk) outputs.
#include "common.cuh"
#include "sort_scan.cuh"

namespace b200 {

int run_batched_nms(const float* boxes, const float* scores, const int32_t* classes, const int32_t* seg_offsets,
                    const int32_t* seg_count, int N, int total_capacity, int num_classes, float iou_thresh,
                    int max_keep, int32_t* keep, int32_t* keep_count, void* workspace, size_t workspace_bytes,
                    bool presorted, int max_slice_hint, cudaStream_t st);

constexpr int kRpnThreads = 1024;
constexpr int kRpnMaxTopk = 16384;   // shared-memory sort capacity (reference configs use 6000 / 12000 / 1000 / 2000)

// ascending key == descending logit; every NaN first (torch's sort treats NaN as the largest value), -0 == +0
__device__ __forceinline__ uint32_t rpn_key(float s) {
  if (s != s) return 0u;
  return desc_key(s);
}

struct RpnSel {
  uint32_t prefix, mask;
  int remaining;
};

__global__ void __launch_bounds__(kRpnThreads)
rpn_topk_filter_kernel(const float* __restrict__ proposals, const float* __restrict__ logits,
                       const int32_t* __restrict__ level_offsets, const float* __restrict__ image_hw, int A, int L,
                       int pre_nms_topk, int cap, float min_size, float* __restrict__ cand_boxes,
                       float* __restrict__ cand_scores, int32_t* __restrict__ cand_lvl,
                       int32_t* __restrict__ seg_offsets, int32_t* __restrict__ cand_count,
                       int32_t* __restrict__ n_invalid) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  unsigned long long* s_keys = reinterpret_cast<unsigned long long*>(s_raw);          // [kRpnMaxTopk]
  int* s_hist = reinterpret_cast<int*>(s_raw + (size_t)kRpnMaxTopk * 8);               // [32][256]
  __shared__ int s_warp[33];
  __shared__ RpnSel s_sel;
  __shared__ int s_count, s_bad;
  const int img = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float img_h = image_hw[2 * img], img_w = image_hw[2 * img + 1];
  const float* lg_img = logits + (size_t)img * A;
  const float* bx_img = proposals + (size_t)img * A * 4;
  const size_t out0 = (size_t)img * cap;
  int running = 0;
  if (threadIdx.x == 0) { s_bad = 0; seg_offsets[img] = img * cap; }

  for (int lvl = 0; lvl < L; ++lvl) {
    const int a0 = level_offsets[lvl], Al = level_offsets[lvl + 1] - a0;
    const int k = min(pre_nms_topk, Al);
    if (k <= 0) continue;                                   // uniform
    const float* lg = lg_img + a0;

    // ---- radix select: the k-th smallest key ------------------------------------------------------------------
    if (threadIdx.x == 0) { s_sel.prefix = 0u; s_sel.mask = 0u; s_sel.remaining = k; s_count = 0; }
    for (int shift = 24; shift >= 0; shift -= 8) {
      for (int i = threadIdx.x; i < 32 * 256; i += kRpnThreads) s_hist[i] = 0;
      __syncthreads();
      const uint32_t prefix = s_sel.prefix, mask = s_sel.mask;
      for (int i = threadIdx.x; i < Al; i += kRpnThreads) {
        const uint32_t key = rpn_key(__ldg(lg + i));
        if ((key & mask) == prefix) atomicAdd(&s_hist[warp * 256 + ((key >> shift) & 255u)], 1);
      }
      __syncthreads();
      if (threadIdx.x < 256) {
        int t = 0;
#pragma unroll 8
        for (int w = 0; w < 32; ++w) t += s_hist[w * 256 + threadIdx.x];
        s_hist[threadIdx.x] = t;                            // row 0 now holds the block histogram
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        int rem = s_sel.remaining, b = 0;
        while (b < 255 && s_hist[b] < rem) { rem -= s_hist[b]; ++b; }
        s_sel.remaining = rem;
        s_sel.prefix = prefix | ((uint32_t)b << shift);
        s_sel.mask = mask | (255u << shift);
      }
      __syncthreads();
    }
    const uint32_t T = s_sel.prefix;
    const int need = s_sel.remaining;                       // how many keys == T belong to the top k (>= 1)
    const int n_less = k - need;

    // ---- collect: keys < T in any order, keys == T in anchor order ----------------------------------------------
    int taken = 0;
    for (int i0 = 0; i0 < Al; i0 += kRpnThreads) {
      const int i = i0 + threadIdx.x;
      const uint32_t key = i < Al ? rpn_key(__ldg(lg + i)) : 0xffffffffu;
      const bool live = i < Al;
      if (live && key < T) s_keys[atomicAdd(&s_count, 1)] = ((unsigned long long)key << 32) | (uint32_t)i;
      const int tie = live && key == T && taken < need;
      if (__syncthreads_or(tie)) {
        int total;
        const int ex = block_exclusive_scan_1024(tie, s_warp, &total);
        if (tie && taken + ex < need) s_keys[n_less + taken + ex] = ((unsigned long long)key << 32) | (uint32_t)i;
        taken += total;
      }
    }
    const int n2 = next_pow2(k);
    for (int i = k + threadIdx.x; i < n2; i += kRpnThreads) s_keys[i] = ~0ull;
    __syncthreads();
    bitonic_sort_u64(s_keys, n2);

    // ---- in sorted order: finite check, clip, min-size test, ordered compaction ---------------------------------
    for (int j0 = 0; j0 < k; j0 += kRpnThreads) {
      const int j = j0 + threadIdx.x;
      int keepf = 0;
      float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
      float s = 0.f;
      if (j < k) {
        const int a = (int)(s_keys[j] & 0xffffffffu);
        b = *reinterpret_cast<const float4*>(bx_img + 4 * (size_t)(a0 + a));
        s = __ldg(lg + a);
        const bool fin = isfinite(b.x) && isfinite(b.y) && isfinite(b.z) && isfinite(b.w) && isfinite(s);
        if (!fin) {
          atomicAdd(&s_bad, 1);
        } else {
          b.x = fminf(fmaxf(b.x, 0.f), img_w); b.y = fminf(fmaxf(b.y, 0.f), img_h);
          b.z = fminf(fmaxf(b.z, 0.f), img_w); b.w = fminf(fmaxf(b.w, 0.f), img_h);
          keepf = (__fsub_rn(b.z, b.x) > min_size) && (__fsub_rn(b.w, b.y) > min_size);
        }
      }
      int total;
      const int ex = block_exclusive_scan_1024(keepf, s_warp, &total);
      // the segment holds `cap` candidates: a raw ABI caller that passes less than sum_l min(pre_nms_topk, A_l) must not
      // make this image overrun its neighbour's segment / the workspace — the excess is dropped and reported
      if (keepf && running + ex < cap) {
        const size_t o = out0 + running + ex;
        *reinterpret_cast<float4*>(cand_boxes + 4 * o) = b;
        cand_scores[o] = s;
        cand_lvl[o] = lvl;
      }
      running += total;
    }
    __syncthreads();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    cand_count[img] = min(running, cap);
    n_invalid[img] = running > cap ? -(running - cap) : s_bad;        // negative: candidates dropped for lack of capacity
  }
}

__global__ void rpn_gather_kernel(const float* __restrict__ cand_boxes, const float* __restrict__ cand_scores,
                                  const int32_t* __restrict__ seg_offsets, const int32_t* __restrict__ keep,
                                  const int32_t* __restrict__ keep_count, int post, float* __restrict__ out_boxes,
                                  float* __restrict__ out_logits) {
  const int img = blockIdx.x;
  const int nk = keep_count[img], base = seg_offsets[img];
  for (int i = threadIdx.x; i < post; i += blockDim.x) {
    const size_t o = (size_t)img * post + i;
    if (i < nk) {
      const int j = base + keep[o];
      *reinterpret_cast<float4*>(out_boxes + 4 * o) = *reinterpret_cast<const float4*>(cand_boxes + 4 * (size_t)j);
      out_logits[o] = cand_scores[j];
    } else {
      *reinterpret_cast<float4*>(out_boxes + 4 * o) = make_float4(0.f, 0.f, 0.f, 0.f);
      out_logits[o] = 0.f;
    }
  }
}

struct RpnWorkspace {
  float* cand_boxes;
  float* cand_scores;
  int32_t* cand_lvl;
  int32_t* seg_offsets;
  int32_t* cand_count;
  int32_t* keep;
  void* nms;
  size_t nms_bytes;
};

static size_t rpn_fixed_bytes(int N, int cap, int post) {
  const size_t tot = (size_t)N * cap;
  return align_up(tot * 16, 256) + 2 * align_up(tot * 4, 256) + 2 * align_up((size_t)N * 4, 256) +
         align_up((size_t)N * post * 4, 256);
}

}  // namespace b200

using namespace b200;

extern "C" size_t b200_rpn_select_workspace_bytes(int N, int cap_per_image, int L, int post_nms_topk) {
  if (N <= 0 || cap_per_image <= 0 || L <= 0 || post_nms_topk < 0) return 256;
  return rpn_fixed_bytes(N, cap_per_image, post_nms_topk) +
         b200_batched_nms_workspace_bytes(N, N * cap_per_image, L);
}

extern "C" int b200_rpn_select_proposals(const float* proposals, const float* logits, const int32_t* level_offsets,
                                         const float* image_hw, int N, int A, int L, int pre_nms_topk,
                                         int post_nms_topk, int cap_per_image, float nms_thresh, float min_box_size,
                                         float* out_boxes, float* out_logits, int32_t* out_count, int32_t* n_invalid,
                                         void* workspace, size_t workspace_bytes, b200_stream_t stream) {
  B200_CHECK_ARG(N >= 0 && A >= 0 && L > 0 && L <= 1024, "rpn_select_proposals: bad shape");
  B200_CHECK_ARG(pre_nms_topk > 0 && pre_nms_topk <= kRpnMaxTopk, "rpn_select_proposals: pre_nms_topk must be in [1, %d]",
                 kRpnMaxTopk);
  B200_CHECK_ARG(post_nms_topk > 0 && cap_per_image > 0 && cap_per_image <= (long long)L * pre_nms_topk,
                 "rpn_select_proposals: bad post_nms_topk / cap_per_image");
  B200_CHECK_ARG((long long)N * cap_per_image < (1ll << 31), "rpn_select_proposals: candidate capacity overflows int32");
  B200_CHECK_ARG(level_offsets && image_hw && out_count && n_invalid && out_boxes && out_logits,
                 "rpn_select_proposals: null tensor");
  B200_CHECK_ARG(A == 0 || (proposals && logits), "rpn_select_proposals: null input");
  B200_CHECK_ARG(((uintptr_t)proposals & 15) == 0, "rpn_select_proposals: proposals must be 16-byte aligned");
  if (N == 0) return B200_OK;
  if (!workspace || workspace_bytes < b200_rpn_select_workspace_bytes(N, cap_per_image, L, post_nms_topk)) {
    set_error("rpn_select_proposals: workspace too small");
    return B200_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const size_t tot = (size_t)N * cap_per_image;
  unsigned char* p = (unsigned char*)workspace;
  RpnWorkspace w;
  w.cand_boxes = (float*)p;    p += align_up(tot * 16, 256);
  w.cand_scores = (float*)p;   p += align_up(tot * 4, 256);
  w.cand_lvl = (int32_t*)p;    p += align_up(tot * 4, 256);
  w.seg_offsets = (int32_t*)p; p += align_up((size_t)N * 4, 256);
  w.cand_count = (int32_t*)p;  p += align_up((size_t)N * 4, 256);
  w.keep = (int32_t*)p;        p += align_up((size_t)N * post_nms_topk * 4, 256);
  w.nms = p;
  w.nms_bytes = workspace_bytes - (size_t)(p - (unsigned char*)workspace);

  const size_t smem = (size_t)kRpnMaxTopk * 8 + 32 * 256 * sizeof(int);
  // set on every call: the attribute is per device and the call is cheap (no process-wide flag to race on)
  B200_CUDA_CALL(cudaFuncSetAttribute(rpn_topk_filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  rpn_topk_filter_kernel<<<N, kRpnThreads, smem, st>>>(proposals, logits, level_offsets, image_hw, A, L, pre_nms_topk,
                                                      cap_per_image, min_box_size, w.cand_boxes, w.cand_scores,
                                                      w.cand_lvl, w.seg_offsets, w.cand_count, n_invalid);
  B200_CUDA_LAUNCH_CHECK("rpn_topk_filter");
  int rc = run_batched_nms(w.cand_boxes, w.cand_scores, w.cand_lvl, w.seg_offsets, w.cand_count, N, (int)tot, L,
                           nms_thresh, post_nms_topk, w.keep, out_count, w.nms, w.nms_bytes, true, pre_nms_topk, st);
  if (rc != B200_OK) return rc;
  rpn_gather_kernel<<<N, 256, 0, st>>>(w.cand_boxes, w.cand_scores, w.seg_offsets, w.keep, out_count, post_nms_topk,
                                      out_boxes, out_logits);
  B200_CUDA_LAUNCH_CHECK("rpn_gather");
  return B200_OK;
}
