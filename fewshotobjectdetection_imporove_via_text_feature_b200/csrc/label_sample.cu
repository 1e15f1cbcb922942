// S1: proposal labelling + sampling for the fine-tune step, one CTA per image.
// Reference: ROIHeads.label_and_sample_proposals / _sample_proposals (defrcn/modeling/roi_heads/roi_heads.py:118-250) ->
// detectron2 pairwise_iou, Matcher(thresholds=[0.5], labels=[0,1], allow_low_quality_matches=False), subsample_labels.
//   match:   per proposal the ground-truth box of highest IoU (first on ties), foreground iff IoU >= threshold
//   sample:  min(max_pos, #fg) foreground + min(batch - that, #bg) background proposals, uniformly at random, foreground
//            rows first (the reference's cat(pos[perm1], neg[perm2]))
// The reference runs this as a Python loop of small torch ops with several host synchronisations per image; here it is
// one launch and the host reads one (n_fg, n_total) pair per image.  Labels and matches are bit-exact with the
// reference (same fp32 IoU expression); the random choice is a counter-based hash of (seed, image, proposal) ordered by
// an in-CTA bitonic sort — same distribution, not torch's RNG stream.
#include "common.cuh"

namespace b200 {

constexpr int kLsThreads = 1024;
constexpr int kLsMaxProps = 4096;      // proposals per image held in the sort buffer
constexpr int kLsMaxGt = 256;          // ground-truth boxes per image staged in shared memory

__device__ __forceinline__ uint32_t ls_hash(uint64_t seed, uint32_t image, uint32_t idx) {
  uint64_t z = seed + ((uint64_t)image << 32 | idx) * 0x9E3779B97F4A7C15ull + 0x632BE59BD9B4E019ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return (uint32_t)((z ^ (z >> 31)) >> 32);
}

// detectron2 pairwise_iou for one pair (fp32, same expression order); best match = first maximum
__device__ __forceinline__ void best_match(const float4 p, const float4* __restrict__ s_gt, int M, float& best, int& arg) {
  const float ap = (p.z - p.x) * (p.w - p.y);
  best = -1.f; arg = 0;
  for (int g = 0; g < M; ++g) {
    const float4 q = s_gt[g];
    const float w = fmaxf(fminf(q.z, p.z) - fmaxf(q.x, p.x), 0.f);
    const float h = fmaxf(fminf(q.w, p.w) - fmaxf(q.y, p.y), 0.f);
    const float inter = w * h;
    const float ag = (q.z - q.x) * (q.w - q.y);
    const float iou = inter > 0.f ? inter / (ag + ap - inter) : 0.f;
    if (iou > best) { best = iou; arg = g; }
  }
  if (M == 0) best = 0.f;
}

__global__ void __launch_bounds__(kLsThreads)
label_sample_kernel(const float4* __restrict__ props, const int32_t* __restrict__ prop_offsets, const float4* __restrict__ gt,
                    const int64_t* __restrict__ gt_classes, const int32_t* __restrict__ gt_offsets, int num_classes,
                    float iou_thresh, int batch, int max_pos, unsigned long long seed,
                    const long long* __restrict__ seed_salt, int append_gt, int pad_background,
                    int32_t* __restrict__ matched_idx,
                    int32_t* __restrict__ matched_label, int32_t* __restrict__ sampled_idx, float4* __restrict__ out_props,
                    int64_t* __restrict__ out_classes, float4* __restrict__ out_gt, int32_t* __restrict__ counts) {
  __shared__ unsigned long long s_key[kLsMaxProps];
  __shared__ float4 s_gt[kLsMaxGt];
  __shared__ int s_nfg;
  const int n = blockIdx.x;
  const int p0 = prop_offsets[n], Pn = prop_offsets[n + 1] - p0;
  const int g0 = gt_offsets[n], M = gt_offsets[n + 1] - g0;
  const int P = Pn + (append_gt ? M : 0);                   // candidates: the proposals, then (optionally) the gt boxes
  const int m0 = p0 + (append_gt ? g0 : 0);                 // first entry of this image in matched_idx / matched_label
  if (P > kLsMaxProps || M > kLsMaxGt || Pn < 0 || M < 0) {
    // the offsets on the device exceed what the host was told (max_props_per_image / max_gt_per_image): the shared
    // arrays are sized for the limits, so refuse the image (no rows) instead of overrunning them
    if (threadIdx.x == 0) { counts[2 * n] = 0; counts[2 * n + 1] = 0; }
    for (int i = threadIdx.x; i < batch; i += kLsThreads) { sampled_idx[(size_t)n * batch + i] = -1; out_classes[(size_t)n * batch + i] = -1; }
    return;
  }
  if (seed_salt) seed += (unsigned long long)*seed_salt;      // device-resident step counter (CUDA-graph replays)
  for (int g = threadIdx.x; g < M; g += kLsThreads) s_gt[g] = gt[g0 + g];
  if (threadIdx.x == 0) s_nfg = 0;
  __syncthreads();
  int cap = 1;
  while (cap < P) cap <<= 1;
  int local_fg = 0;
  for (int i = threadIdx.x; i < cap; i += kLsThreads) {
    unsigned long long key = ~0ull;
    if (i < P) {
      float best;
      int arg;
      best_match(i < Pn ? props[p0 + i] : s_gt[i - Pn], s_gt, M, best, arg);
      const bool fg = M > 0 && best >= iou_thresh;
      if (matched_idx) matched_idx[m0 + i] = arg;
      if (matched_label) matched_label[m0 + i] = fg ? 1 : 0;
      local_fg += fg;
      // [63] background flag, [62:31] random key, [30:0] proposal index: foreground first, random order within each class
      key = ((unsigned long long)(fg ? 0 : 1) << 63) | ((unsigned long long)ls_hash(seed, n, i) << 31) | (unsigned long long)i;
    }
    s_key[i] = key;
  }
  if (local_fg) atomicAdd(&s_nfg, local_fg);
  __syncthreads();
  // bitonic sort, ascending
  for (int k = 2; k <= cap; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < cap; i += kLsThreads) {
        const int l = i ^ j;
        if (l > i) {
          const unsigned long long a = s_key[i], b = s_key[l];
          const bool up = (i & k) == 0;
          if ((a > b) == up) { s_key[i] = b; s_key[l] = a; }
        }
      }
      __syncthreads();
    }
  }
  const int nfg = s_nfg, nbg = P - nfg;
  const int num_pos = min(max_pos, nfg), num_neg = min(batch - num_pos, nbg);
  const int total = num_pos + num_neg;
  if (threadIdx.x == 0) { counts[2 * n] = num_pos; counts[2 * n + 1] = total; }
  for (int j = threadIdx.x; j < batch; j += kLsThreads) {
    const size_t o = (size_t)n * batch + j;
    if (j >= total) {
      sampled_idx[o] = -1;
      out_props[o] = make_float4(0.f, 0.f, 0.f, 0.f);
      out_gt[o] = make_float4(0.f, 0.f, 0.f, 0.f);
      out_classes[o] = pad_background ? (int64_t)num_classes : -1;
      continue;
    }
    const unsigned long long key = s_key[j < num_pos ? j : nfg + (j - num_pos)];
    const int i = (int)(key & 0x7fffffffull);
    const float4 p = i < Pn ? props[p0 + i] : s_gt[i - Pn];
    float best;
    int arg;
    best_match(p, s_gt, M, best, arg);
    const bool fg = M > 0 && best >= iou_thresh;
    sampled_idx[o] = i;
    out_props[o] = p;
    out_classes[o] = fg ? gt_classes[g0 + arg] : (int64_t)num_classes;
    out_gt[o] = M > 0 ? s_gt[arg] : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

}  // namespace b200

using namespace b200;

extern "C" int b200_label_sample_proposals(const float* proposals, const int32_t* prop_offsets, const float* gt_boxes,
                                           const int64_t* gt_classes, const int32_t* gt_offsets, int num_images,
                                           int max_props_per_image, int max_gt_per_image, int num_classes, float iou_thresh,
                                           int batch_per_image, int max_positive, unsigned long long seed,
                                           const int64_t* seed_salt, int append_gt, int pad_background, int32_t* matched_idx, int32_t* matched_label, int32_t* sampled_idx,
                                           float* out_proposals, int64_t* out_classes, float* out_gt_boxes, int32_t* counts,
                                           b200_stream_t stream) {
  B200_CHECK_ARG(prop_offsets && gt_offsets && sampled_idx && out_proposals && out_classes && out_gt_boxes && counts,
                 "label_sample_proposals: null tensor");
  B200_CHECK_ARG(num_images >= 0 && batch_per_image > 0 && max_positive >= 0 && max_positive <= batch_per_image,
                 "label_sample_proposals: bad sizes");
  if (max_props_per_image + (append_gt ? max_gt_per_image : 0) > kLsMaxProps || max_gt_per_image > kLsMaxGt) {
    set_error("label_sample_proposals: at most %d proposals and %d ground-truth boxes per image", kLsMaxProps, kLsMaxGt);
    return B200_ERR_UNSUPPORTED;
  }
  B200_CHECK_ARG((((uintptr_t)proposals | (uintptr_t)gt_boxes | (uintptr_t)out_proposals | (uintptr_t)out_gt_boxes) & 15) == 0,
                 "label_sample_proposals: box tensors must be 16-byte aligned");
  if (num_images == 0) return B200_OK;
  label_sample_kernel<<<num_images, kLsThreads, 0, (cudaStream_t)stream>>>(
      (const float4*)proposals, prop_offsets, (const float4*)gt_boxes, gt_classes, gt_offsets, num_classes, iou_thresh,
      batch_per_image, max_positive, seed, (const long long*)seed_salt, append_gt, pad_background, matched_idx, matched_label, sampled_idx, (float4*)out_proposals, out_classes,
      (float4*)out_gt_boxes, counts);
  B200_CUDA_LAUNCH_CHECK("label_sample_proposals");
  return B200_OK;
}
