/*
 * oracle/oracle_ref.c — TEST INFRASTRUCTURE ONLY (never imported by the product path).
 *
 * Plain-C, single-threaded CPU restatement of the third-party arithmetic the reference's
 * ROI-head hot path executes.  The reference itself is 100 % Python; these primitives live
 * in un-vendored wheels pinned by /root/reference/requirements.txt:13,75,77:
 *
 *   - torchvision==0.8.1  ops.roi_align   (CPU kernel ROIAlign_cpu / roi_align_kernel.cpp;
 *                                          same algorithm as the installed 0.26 kernel)
 *                         ops.nms         (nms_cpu_kernel)
 *                         ops.batched_nms (coordinate-offset trick, boxes.py)
 *   - detectron2==0.3     Box2BoxTransform.apply_deltas, Boxes.clip, batched_nms wrapper
 *
 * Call sites in the reference that reach them:
 *   roi_align      : defrcn/modeling/roi_heads/roi_heads.py:300-305,339-344 ; calibration_layer.py:27,100
 *   apply_deltas   : defrcn/modeling/roi_heads/fast_rcnn.py:306-324
 *   clip/threshold : defrcn/modeling/roi_heads/fast_rcnn.py:104-122
 *   batched_nms    : defrcn/modeling/roi_heads/fast_rcnn.py:125-128
 *
 * Parity pinning: the reference has no tests.  This file is pinned against
 * (a) the installed torchvision 0.26 CPU ops (tests/test_oracle_pinning.py) and
 * (b) golden vectors produced by importing the reference's own fast_rcnn.py unchanged
 *     (oracle/gen_golden.py -> tests/golden/).
 *
 * Compile with:  gcc -O2 -ffp-contract=off -fno-fast-math -shared -fPIC   (see oracle/Makefile)
 * -ffp-contract=off matters: the published kernels are compiled without FMA contraction on x86.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------ */
/* ROIAlign forward — torchvision roi_align (aligned flag, adaptive sampling_ratio<=0)    */
/* layout: input NCHW fp32, rois (R,5) = [batch_idx, x1, y1, x2, y2], output (R,C,PH,PW)  */
/* ------------------------------------------------------------------------------------ */
typedef struct {
  int pos1, pos2, pos3, pos4;
  float w1, w2, w3, w4;
} precalc_t;

static void pre_calc_bilinear(int height, int width, int pooled_h, int pooled_w,
                              float roi_start_h, float roi_start_w, float bin_size_h,
                              float bin_size_w, int grid_h, int grid_w, precalc_t* pc) {
  int idx = 0;
  for (int ph = 0; ph < pooled_h; ph++) {
    for (int pw = 0; pw < pooled_w; pw++) {
      for (int iy = 0; iy < grid_h; iy++) {
        const float yy = roi_start_h + ph * bin_size_h +
                         ((float)iy + .5f) * bin_size_h / (float)grid_h;
        for (int ix = 0; ix < grid_w; ix++) {
          const float xx = roi_start_w + pw * bin_size_w +
                           ((float)ix + .5f) * bin_size_w / (float)grid_w;
          float x = xx, y = yy;
          if (y < -1.0f || y > (float)height || x < -1.0f || x > (float)width) {
            precalc_t z = {0, 0, 0, 0, 0.f, 0.f, 0.f, 0.f};
            pc[idx++] = z;
            continue;
          }
          if (y <= 0) y = 0;
          if (x <= 0) x = 0;
          int y_low = (int)y, x_low = (int)x, y_high, x_high;
          if (y_low >= height - 1) { y_high = y_low = height - 1; y = (float)y_low; }
          else y_high = y_low + 1;
          if (x_low >= width - 1) { x_high = x_low = width - 1; x = (float)x_low; }
          else x_high = x_low + 1;
          const float ly = y - y_low, lx = x - x_low, hy = 1.f - ly, hx = 1.f - lx;
          precalc_t p;
          p.w1 = hy * hx; p.w2 = hy * lx; p.w3 = ly * hx; p.w4 = ly * lx;
          p.pos1 = y_low * width + x_low;  p.pos2 = y_low * width + x_high;
          p.pos3 = y_high * width + x_low; p.pos4 = y_high * width + x_high;
          pc[idx++] = p;
        }
      }
    }
  }
}

int oracle_roi_align_fwd(const float* input, const float* rois, int N, int C, int H, int W,
                         int R, int PH, int PW, float spatial_scale, int sampling_ratio,
                         int aligned, float* output) {
  (void)N;
  for (int n = 0; n < R; n++) {
    const float* roi = rois + 5 * n;
    const int b = (int)roi[0];
    const float offset = aligned ? 0.5f : 0.0f;
    const float roi_start_w = roi[1] * spatial_scale - offset;
    const float roi_start_h = roi[2] * spatial_scale - offset;
    const float roi_end_w = roi[3] * spatial_scale - offset;
    const float roi_end_h = roi[4] * spatial_scale - offset;
    float roi_w = roi_end_w - roi_start_w, roi_h = roi_end_h - roi_start_h;
    if (!aligned) { roi_w = fmaxf(roi_w, 1.f); roi_h = fmaxf(roi_h, 1.f); }
    const float bin_h = roi_h / (float)PH, bin_w = roi_w / (float)PW;
    const int gh = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(roi_h / PH);
    const int gw = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(roi_w / PW);
    const int ghc = gh > 0 ? gh : 0, gwc = gw > 0 ? gw : 0;
    const float count = (float)(ghc * gwc > 1 ? ghc * gwc : 1);
    const size_t npc = (size_t)ghc * gwc * PH * PW;
    precalc_t* pc = (precalc_t*)malloc((npc ? npc : 1) * sizeof(precalc_t));
    if (!pc) return -1;
    pre_calc_bilinear(H, W, PH, PW, roi_start_h, roi_start_w, bin_h, bin_w, ghc, gwc, pc);
    for (int c = 0; c < C; c++) {
      const float* in = input + ((size_t)b * C + c) * H * W;
      float* out = output + ((size_t)n * C + c) * PH * PW;
      size_t idx = 0;
      for (int ph = 0; ph < PH; ph++)
        for (int pw = 0; pw < PW; pw++) {
          float acc = 0.f;
          for (int iy = 0; iy < ghc; iy++)
            for (int ix = 0; ix < gwc; ix++) {
              const precalc_t p = pc[idx++];
              acc += p.w1 * in[p.pos1] + p.w2 * in[p.pos2] + p.w3 * in[p.pos3] +
                     p.w4 * in[p.pos4];
            }
          out[ph * PW + pw] = acc / count;
        }
    }
    free(pc);
  }
  return 0;
}

/* ROIAlign backward — torchvision roi_align_backward CPU kernel (serial adds; the CUDA   */
/* kernel uses atomicAdd, so summation order there is non-deterministic).                 */
int oracle_roi_align_bwd(const float* grad_out, const float* rois, int N, int C, int H, int W,
                         int R, int PH, int PW, float spatial_scale, int sampling_ratio,
                         int aligned, float* grad_in /* zero-initialised by caller */) {
  (void)N;
  for (int n = 0; n < R; n++) {
    const float* roi = rois + 5 * n;
    const int b = (int)roi[0];
    const float offset = aligned ? 0.5f : 0.0f;
    const float roi_start_w = roi[1] * spatial_scale - offset;
    const float roi_start_h = roi[2] * spatial_scale - offset;
    const float roi_end_w = roi[3] * spatial_scale - offset;
    const float roi_end_h = roi[4] * spatial_scale - offset;
    float roi_w = roi_end_w - roi_start_w, roi_h = roi_end_h - roi_start_h;
    if (!aligned) { roi_w = fmaxf(roi_w, 1.f); roi_h = fmaxf(roi_h, 1.f); }
    const float bin_h = roi_h / (float)PH, bin_w = roi_w / (float)PW;
    const int gh = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(roi_h / PH);
    const int gw = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(roi_w / PW);
    const float count = (float)(gh * gw);
    for (int c = 0; c < C; c++) {
      float* gin = grad_in + ((size_t)b * C + c) * H * W;
      const float* go = grad_out + ((size_t)n * C + c) * PH * PW;
      for (int ph = 0; ph < PH; ph++)
        for (int pw = 0; pw < PW; pw++) {
          const float g = go[ph * PW + pw];
          for (int iy = 0; iy < gh; iy++) {
            const float yy = roi_start_h + ph * bin_h + ((float)iy + .5f) * bin_h / (float)gh;
            for (int ix = 0; ix < gw; ix++) {
              const float xx = roi_start_w + pw * bin_w + ((float)ix + .5f) * bin_w / (float)gw;
              float x = xx, y = yy;
              if (y < -1.0f || y > (float)H || x < -1.0f || x > (float)W) continue;
              if (y <= 0) y = 0;
              if (x <= 0) x = 0;
              int y_low = (int)y, x_low = (int)x, y_high, x_high;
              if (y_low >= H - 1) { y_high = y_low = H - 1; y = (float)y_low; } else y_high = y_low + 1;
              if (x_low >= W - 1) { x_high = x_low = W - 1; x = (float)x_low; } else x_high = x_low + 1;
              const float ly = y - y_low, lx = x - x_low, hy = 1.f - ly, hx = 1.f - lx;
              const float w1 = hy * hx, w2 = hy * lx, w3 = ly * hx, w4 = ly * lx;
              gin[y_low * W + x_low] += g * w1 / count;
              gin[y_low * W + x_high] += g * w2 / count;
              gin[y_high * W + x_low] += g * w3 / count;
              gin[y_high * W + x_high] += g * w4 / count;
            }
          }
        }
    }
  }
  return 0;
}

/* ------------------------------------------------------------------------------------ */
/* Box2BoxTransform.apply_deltas (detectron2 0.3 box_regression.py) + Boxes.clip          */
/* deltas (R, K*4), proposals (R,4) xyxy; out (R, K*4).  fp32, no FMA contraction.        */
/* NB: expf() is this libm's; torch CPU uses Sleef — results may differ in the last ulp,  */
/* which is why box values are compared with a tolerance while NMS decisions are checked  */
/* bit-exactly on identical box inputs.                                                   */
/* ------------------------------------------------------------------------------------ */
int oracle_apply_deltas(const float* deltas, const float* props, int R, int K, float wx,
                        float wy, float ww, float wh, float scale_clamp, float* out) {
  for (int r = 0; r < R; r++) {
    const float* b = props + 4 * r;
    const float widths = b[2] - b[0], heights = b[3] - b[1];
    const float ctr_x = b[0] + 0.5f * widths, ctr_y = b[1] + 0.5f * heights;
    for (int k = 0; k < K; k++) {
      const float* d = deltas + ((size_t)r * K + k) * 4;
      float dx = d[0] / wx, dy = d[1] / wy, dw = d[2] / ww, dh = d[3] / wh;
      if (dw > scale_clamp) dw = scale_clamp;
      if (dh > scale_clamp) dh = scale_clamp;
      const float pcx = dx * widths + ctr_x, pcy = dy * heights + ctr_y;
      const float pw = expf(dw) * widths, ph = expf(dh) * heights;
      float* o = out + ((size_t)r * K + k) * 4;
      o[0] = pcx - 0.5f * pw; o[1] = pcy - 0.5f * ph;
      o[2] = pcx + 0.5f * pw; o[3] = pcy + 0.5f * ph;
    }
  }
  return 0;
}

void oracle_clip_boxes(float* boxes, int n, float h, float w) {
  for (int i = 0; i < n; i++) {
    float* b = boxes + 4 * i;
    b[0] = fminf(fmaxf(b[0], 0.f), w); b[1] = fminf(fmaxf(b[1], 0.f), h);
    b[2] = fminf(fmaxf(b[2], 0.f), w); b[3] = fminf(fmaxf(b[3], 0.f), h);
  }
}

/* ------------------------------------------------------------------------------------ */
/* torchvision nms (CPU kernel): stable descending sort by score, greedy suppression with */
/* iou = inter / (area_i + area_j - inter), strict `>` threshold.                          */
/* ------------------------------------------------------------------------------------ */
typedef struct { float s; int64_t i; } sidx_t;
static int cmp_desc_stable(const void* a, const void* b) {
  const sidx_t* x = (const sidx_t*)a; const sidx_t* y = (const sidx_t*)b;
  if (x->s > y->s) return -1;
  if (x->s < y->s) return 1;
  return (x->i < y->i) ? -1 : (x->i > y->i);
}

int64_t oracle_nms(const float* boxes, const float* scores, int64_t n, float thr,
                   int64_t* keep) {
  if (n == 0) return 0;
  sidx_t* ord = (sidx_t*)malloc(n * sizeof(sidx_t));
  float* areas = (float*)malloc(n * sizeof(float));
  uint8_t* sup = (uint8_t*)calloc(n, 1);
  for (int64_t i = 0; i < n; i++) {
    ord[i].s = scores[i]; ord[i].i = i;
    areas[i] = (boxes[4 * i + 2] - boxes[4 * i]) * (boxes[4 * i + 3] - boxes[4 * i + 1]);
  }
  qsort(ord, n, sizeof(sidx_t), cmp_desc_stable);
  int64_t nk = 0;
  for (int64_t _i = 0; _i < n; _i++) {
    const int64_t i = ord[_i].i;
    if (sup[i]) continue;
    keep[nk++] = i;
    const float ix1 = boxes[4 * i], iy1 = boxes[4 * i + 1], ix2 = boxes[4 * i + 2],
                iy2 = boxes[4 * i + 3], iarea = areas[i];
    for (int64_t _j = _i + 1; _j < n; _j++) {
      const int64_t j = ord[_j].i;
      if (sup[j]) continue;
      const float xx1 = fmaxf(ix1, boxes[4 * j]), yy1 = fmaxf(iy1, boxes[4 * j + 1]);
      const float xx2 = fminf(ix2, boxes[4 * j + 2]), yy2 = fminf(iy2, boxes[4 * j + 3]);
      const float w = fmaxf(0.f, xx2 - xx1), h = fmaxf(0.f, yy2 - yy1);
      const float inter = w * h;
      const float ovr = inter / (iarea + areas[j] - inter);
      if (ovr > thr) sup[j] = 1;
    }
  }
  free(ord); free(areas); free(sup);
  return nk;
}

/* torchvision batched_nms coordinate trick (what detectron2 0.3 runs for < 40000 boxes):  */
/*   offsets = idxs.to(boxes) * (boxes.max() + 1); nms(boxes + offsets[:,None], scores)   */
int64_t oracle_batched_nms(const float* boxes, const float* scores, const int64_t* idxs,
                           int64_t n, float thr, int64_t* keep) {
  if (n == 0) return 0;
  float mx = boxes[0];
  for (int64_t i = 1; i < 4 * n; i++) mx = boxes[i] > mx ? boxes[i] : mx;
  const float m1 = mx + 1.0f;
  float* sh = (float*)malloc(4 * n * sizeof(float));
  for (int64_t i = 0; i < n; i++) {
    const float off = (float)idxs[i] * m1;
    for (int k = 0; k < 4; k++) sh[4 * i + k] = boxes[4 * i + k] + off;
  }
  const int64_t nk = oracle_nms(sh, scores, n, thr, keep);
  free(sh);
  return nk;
}

/* ------------------------------------------------------------------------------------ */
/* fast_rcnn_inference_single_image (fast_rcnn.py:90-134), given per-class boxes (R,K*4)  */
/* and probabilities (R,K+1).  Outputs at most `topk` rows; returns number written, and   */
/* *n_candidates = number of (roi,class) pairs with score > thresh (the compaction count).*/
/* ------------------------------------------------------------------------------------ */
int64_t oracle_fast_rcnn_inference_single_image(
    const float* boxes_in, const float* probs, int R, int K, float img_h, float img_w,
    float score_thresh, float nms_thresh, int64_t topk, float* out_boxes, float* out_scores,
    int64_t* out_classes, int64_t* out_roi_inds, int64_t* n_candidates,
    int64_t* cand_inds /* optional (R*K,2) */) {
  float* boxes = (float*)malloc((size_t)R * K * 4 * sizeof(float));
  memcpy(boxes, boxes_in, (size_t)R * K * 4 * sizeof(float));
  oracle_clip_boxes(boxes, R * K, img_h, img_w);
  int64_t n = 0;
  float* cb = (float*)malloc((size_t)R * K * 4 * sizeof(float) + 16);
  float* cs = (float*)malloc((size_t)R * K * sizeof(float) + 16);
  int64_t* cc = (int64_t*)malloc((size_t)R * K * sizeof(int64_t) + 16);
  int64_t* cr = (int64_t*)malloc((size_t)R * K * sizeof(int64_t) + 16);
  for (int r = 0; r < R; r++)
    for (int k = 0; k < K; k++) {
      const float s = probs[(size_t)r * (K + 1) + k];
      if (s > score_thresh) {
        memcpy(cb + 4 * n, boxes + ((size_t)r * K + k) * 4, 4 * sizeof(float));
        cs[n] = s; cc[n] = k; cr[n] = r;
        if (cand_inds) { cand_inds[2 * n] = r; cand_inds[2 * n + 1] = k; }
        n++;
      }
    }
  *n_candidates = n;
  int64_t* keep = (int64_t*)malloc((n + 1) * sizeof(int64_t));
  int64_t nk = oracle_batched_nms(cb, cs, cc, n, nms_thresh, keep);
  if (topk >= 0 && nk > topk) nk = topk;
  for (int64_t i = 0; i < nk; i++) {
    const int64_t j = keep[i];
    memcpy(out_boxes + 4 * i, cb + 4 * j, 4 * sizeof(float));
    out_scores[i] = cs[j]; out_classes[i] = cc[j]; out_roi_inds[i] = cr[j];
  }
  free(boxes); free(cb); free(cs); free(cc); free(cr); free(keep);
  return nk;
}
