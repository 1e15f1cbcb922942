/*
 * b200roi.h — C ABI of the B200-native (sm_100a) ROI-head hot path.
 *
 * Drop-in boundary for the DeFRCN text-fused ROI head of
 * hoangpnhat/FewShotObjectDetection_imporove_via_text_feature.  The reference is pure Python and binds
 * no native code of its own; each entry point below replaces the third-party / ATen call the reference
 * makes at the cited file:line (paths relative to the reference root).  INTEGRATION.md shows the
 * ctypes stubs a maintainer adds on the reference side.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless named host_*; no allocation happens inside the library:
 *     the caller passes outputs and (where needed) a workspace sized by the matching *_workspace_bytes();
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*) and re-entrant per stream.  The only
 *     process-wide state is the set of implementation switches of b200_set_option (A/B measurement and tests; every
 *     setting computes the same results) and a cached driver entry point; neither is written after start-up by the
 *     compute calls themselves;
 *   - return value: B200_OK or a negative B200_ERR_*; never throws.  b200_last_error() returns a
 *     thread-local description of the last failure on the calling thread;
 *   - dtype: B200_F32 | B200_BF16 (storage type of the feature maps / activations; accumulation is fp32);
 *   - layout: B200_NCHW (torch contiguous) | B200_NHWC (torch channels_last) for 4-d tensors.
 */
#ifndef B200ROI_H_
#define B200ROI_H_

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define B200_API __attribute__((visibility("default")))
#else
#define B200_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef void* b200_stream_t; /* cudaStream_t */

enum { B200_OK = 0, B200_ERR_INVALID = -1, B200_ERR_CUDA = -2, B200_ERR_WORKSPACE = -3, B200_ERR_UNSUPPORTED = -4 };
enum { B200_F32 = 0, B200_BF16 = 1 };
enum { B200_NCHW = 0, B200_NHWC = 1 };

B200_API int b200_abi_version(void);
B200_API const char* b200_last_error(void);
/* process-wide implementation switches (for A/B measurement; results are parity-tested under every setting):
 *   "roi_align_bf16_impl": 0 = CUDA-core per-bin-window kernel, 1 = per-ROI TMA + ldmatrix + mma.sync kernel,
 *                          2 = slice-resident TMA + ldmatrix + mma.sync kernel (default; needs roi_batch_offsets)
 *   "roi_align_bwd_impl":  bf16 NHWC 7x7 path — 0 = fp32 weight-table row kernel, 1 = per-pixel CSR gather,
 *                          2 = 4x4-pixel-tile gather on mma.sync (default; C % 64 == 0, else 1).  The format of a
 *                          b200_roi_align_bwd_plan buffer follows the setting at plan time; b200_roi_align_bwd_planned
 *                          must be called under the same setting
 *   "roi_bwd_tile_variant": 0 (default: tiles launched heaviest first) / 1 / 2 (pipelining variants) / 3 (the default's
 *                          kernel in map order) of the tile gather, same bits */
B200_API int b200_set_option(const char* key, int value);

/* ---------------------------------------------------------------------------------------------------
 * G1 + G2  Gradient Decoupled Layer + AffineLayer
 *   replaces defrcn/modeling/meta_arch/gdl.py:6-38 as called at defrcn/modeling/meta_arch/rcnn.py:94-97
 *   fwd:  y = x * weight[c] + bias[c]           (GDL is the identity in forward)
 *   bwd:  grad_x = grad_y * weight[c] * lambda  (GDL scales by lambda: cfg.MODEL.ROI_HEADS.BACKWARD_SCALE)
 *         grad_w[c] = sum_{n,h,w} grad_y * x ;  grad_b[c] = sum_{n,h,w} grad_y   (deterministic two-pass)
 *   x / grad_x use (in_dtype, in_layout); y / grad_y use (out_dtype, out_layout).  weight==NULL means
 *   identity scale (pure layout / dtype conversion), bias may be NULL.
 * ------------------------------------------------------------------------------------------------- */
B200_API int b200_gdl_affine_fwd(const void* x, const float* weight, const float* bias, void* y, int N, int C, int H,
                        int W, int in_dtype, int in_layout, int out_dtype, int out_layout, b200_stream_t stream);
B200_API size_t b200_gdl_affine_bwd_workspace_bytes(int N, int C, int H, int W);
B200_API int b200_gdl_affine_bwd(const void* grad_y, const void* x, const float* weight, float lambda, void* grad_x,
                        float* grad_w, float* grad_b, int N, int C, int H, int W, int in_dtype, int in_layout,
                        int out_dtype, int out_layout, void* workspace, size_t workspace_bytes,
                        b200_stream_t stream);

/* ---------------------------------------------------------------------------------------------------
 * P1 / P1b / Q1  ROIAlign (detectron2 ROIPooler "ROIAlignV2" -> torchvision.ops.roi_align)
 *   replaces the call at defrcn/modeling/roi_heads/roi_heads.py:300-305,339-340 and
 *   defrcn/evaluation/calibration_layer.py:27,100.
 *   feat (N,C,H,W) in (dtype,in_layout); rois (R,5) fp32 [batch_idx,x1,y1,x2,y2] in image coordinates;
 *   out (R,C,PHO,PWO) in (dtype,out_layout), PHO = ceil(pooled_h / bin_step).  sampling_ratio<=0 is the adaptive
 *   ceil(roi/pooled) grid.
 *   bin_step (>=1): compute and store only the bins (0, step, 2*step, ...) of each axis.  bin_step = 2 with 7x7
 *   pooling is what res5 actually consumes: its first block reads the pooled map through 1x1 stride-2 convolutions
 *   (roi_heads.py:313-337, RESNETS.STRIDE_IN_1X1=True), so 33 of the 49 bins are never read.  1 = the reference op.
 *   roi_batch_offsets (N+1 int32, device; may be NULL in fwd): prefix of per-image ROI counts when the ROIs are
 *   grouped by image (as detectron2's convert_boxes_to_pooler_format produces them).  Given it, bf16 channels-last
 *   7x7 pooling runs the slice-resident tensor-core kernel (csrc/roi_align_slice.cu); without it, the per-ROI kernel.
 *   The workspace holds an NHWC copy of the map when in_layout==NCHW plus, for bf16, R geometry records.
 *   bwd is atomic-free and deterministic: grad_feat is fully overwritten (no pre-zeroing needed).
 * ------------------------------------------------------------------------------------------------- */
B200_API size_t b200_roi_align_fwd_workspace_bytes(int N, int C, int H, int W, int R, int dtype, int in_layout);
B200_API int b200_roi_align_fwd(const void* feat, const float* rois, const int32_t* roi_batch_offsets, void* out, int N,
                       int C, int H, int W, int R, int pooled_h, int pooled_w, int bin_step, float spatial_scale,
                       int sampling_ratio, int aligned, int dtype, int in_layout, int out_layout, void* workspace,
                       size_t workspace_bytes, b200_stream_t stream);
B200_API size_t b200_roi_align_bwd_workspace_bytes(int N, int C, int H, int W, int R, int pooled_h, int pooled_w,
                                          int bin_step, int dtype, int grad_in_layout, int grad_out_layout);
B200_API int b200_roi_align_bwd(const void* grad_out, const float* rois, const int32_t* roi_batch_offsets,
                       void* grad_feat, int N, int C, int H, int W, int R, int pooled_h, int pooled_w, int bin_step,
                       float spatial_scale, int sampling_ratio, int aligned, int dtype, int grad_out_layout,
                       int grad_in_layout, void* workspace, size_t workspace_bytes, b200_stream_t stream);
/* Split form of b200_roi_align_bwd for bf16 channels-last gradients (7x7 pooling): the per-pixel gather lists depend
 * only on the ROIs, so `_plan` can run ahead of the backward pass (e.g. on a side stream during the forward) and
 * `_planned` is then a single gather launch.  `_plan_bytes` returns 0 when the shape is not covered. */
B200_API size_t b200_roi_align_bwd_plan_bytes(int N, int C, int H, int W, int R, int pooled_h, int pooled_w, int bin_step);
B200_API int b200_roi_align_bwd_plan(const float* rois, const int32_t* roi_batch_offsets, int N, int C, int H, int W, int R,
                            int pooled_h, int pooled_w, int bin_step, float spatial_scale, int sampling_ratio, int aligned,
                            void* plan, size_t plan_bytes, b200_stream_t stream);
B200_API int b200_roi_align_bwd_planned(const void* grad_out, const void* plan, size_t plan_bytes, void* grad_feat, int N, int C,
                               int H, int W, int R, int pooled_h, int pooled_w, int bin_step, b200_stream_t stream);


/* ---------------------------------------------------------------------------------------------------
 * D1 + D2 + D3 (front)  softmax, Box2BoxTransform.apply_deltas, Boxes.clip, score threshold, ordered
 * compaction.  Replaces defrcn/modeling/roi_heads/fast_rcnn.py:306-334 and :104-122.
 *   scores_in (R,K+1): logits (input_is_prob=0, softmax applied) or probabilities (input_is_prob=1);
 *   deltas (R,4K) or (R,4) when cls_agnostic; proposals (R,4); roi_offsets (N+1) int32 prefix of the
 *   per-image ROI counts (device); image_hw (N,2) fp32 (h,w) (device).
 *   probs_out (R,K+1) may be NULL.  Candidates of image i are written, in torch.nonzero() order
 *   (roi-major, class-minor), at [roi_offsets[i]*K, roi_offsets[i]*K + cand_count[i]) of
 *   cand_boxes (R*K,4) / cand_scores / cand_roi (index within the image) / cand_cls.
 * -------------------------------------------------------------------------------------------------  *   workspace (optional, b200_softmax_decode_compact_workspace_bytes; max_rois_per_image >= every image's ROI count):
 *   spreads an image over ceil(max_rois_per_image / 256) CTAs in two launches (count, write) — same output order and
 *   bits; without it one CTA walks an image's rows (fine up to ~1000 proposals per image).
 */
B200_API size_t b200_softmax_decode_compact_workspace_bytes(int N, int max_rois_per_image);
B200_API int b200_softmax_decode_compact(const float* scores_in, int input_is_prob, const float* deltas,
                                const float* proposals, const int32_t* roi_offsets, const float* image_hw,
                                int N, int R, int K, int cls_agnostic, float wx, float wy, float ww, float wh,
                                float score_thresh, float* probs_out, float* cand_boxes, float* cand_scores,
                                int32_t* cand_roi, int32_t* cand_cls, int32_t* cand_count, int max_rois_per_image, void* workspace,
                                size_t workspace_bytes,
                                b200_stream_t stream);

/* ---------------------------------------------------------------------------------------------------
 * D3 (back)  per-class NMS == detectron2.layers.batched_nms (torchvision coordinate-offset trick) followed
 * by keep[:topk].  Replaces defrcn/modeling/roi_heads/fast_rcnn.py:125-128.
 *   Segment i (one image) holds seg_count[i] boxes starting at seg_offsets[i] (device arrays; offsets are
 *   the capacities' prefix so segments may be sparse).  keep (N,max_keep) receives indices relative to the
 *   segment start, ordered by descending score (ties: ascending index); keep_count (N).
 *   Bit-exact against the reference path for segments below 40000 boxes (above that detectron2 0.3
 *   switches to an un-offset per-class loop, which this kernel follows as well).
 *   max_class_slice: an upper bound the caller knows for the boxes of ONE class in ONE segment (e.g. the image's ROI
 *   count: a ROI yields at most one candidate per class), 0 = unknown (total_capacity is assumed).  It only decides
 *   whether the launch for class slices above 4096 boxes (1024 threads, 197 KB shared memory) is issued at all.
 * ------------------------------------------------------------------------------------------------- */
B200_API size_t b200_batched_nms_workspace_bytes(int N, int total_capacity, int num_classes);
B200_API int b200_batched_nms(const float* boxes, const float* scores, const int32_t* classes,
                     const int32_t* seg_offsets, const int32_t* seg_count, int N, int total_capacity,
                     int num_classes, float iou_thresh, int max_keep, int max_class_slice, int32_t* keep,
                     int32_t* keep_count, void* workspace, size_t workspace_bytes, b200_stream_t stream);

/* gather kept candidates into padded detection tensors: boxes (N,max_keep,4), scores (N,max_keep),
 * classes / roi_inds (N,max_keep) int64 — fast_rcnn.py:128-134 */
B200_API int b200_gather_detections(const float* cand_boxes, const float* cand_scores, const int32_t* cand_roi,
                           const int32_t* cand_cls, const int32_t* seg_offsets, const int32_t* keep,
                           const int32_t* keep_count, int N, int max_keep, float* out_boxes,
                           float* out_scores, int64_t* out_classes, int64_t* out_roi_inds,
                           b200_stream_t stream);

/* ---------------------------------------------------------------------------------------------------
 * Q2  PCB cosine calibration.  Replaces defrcn/evaluation/calibration_layer.py:110-123 (per-detection
 * sklearn cosine_similarity on the CPU).  scores (n) sorted descending, updated in place for detections
 * ileft <= i < iright where ileft = #(score > upper), iright = #(score > lower), unless
 * exclude[class]!=0:  s = alpha*s + (1-alpha)*cos(feats[i], prototypes[class]).
 * feats (n,D) holds features for ALL n detections (row i <-> detection i).
 * ------------------------------------------------------------------------------------------------- */
B200_API int b200_pcb_cosine_blend(float* scores, const float* feats, const float* prototypes, const int64_t* classes,
                          const uint8_t* exclude, int n, int D, int K, float alpha, float lower, float upper,
                          b200_stream_t stream);

/* ---------------------------------------------------------------------------------------------------
 * A1..A6, C1, C1'  text-fusion chain on tcgen05 tensor cores.
 * b200_gemm_bf16: D[M,N] = act(A[M,K] * B[N,K]^T + bias[N])   A,B bf16 row-major (K contiguous), fp32
 *   accumulation in TMEM, TMA-fed 128B-swizzled smem tiles.  Replaces the nn.Linear calls at
 *   defrcn/modeling/roi_heads/attentive_modules.py:124,166-175,72 and fast_rcnn.py:407,415.
 *   M,N arbitrary (TMA clips); K % 8 == 0; lda/ldb/ldd in elements, multiples of 8.
 *   out_dtype B200_F32|B200_BF16; relu: 0/1.  d2/ldd2 (optional, may be NULL): a second bf16 copy of the
 *   result, used to feed the next GEMM while the fp32 copy is returned to the caller.
 * ------------------------------------------------------------------------------------------------- */
B200_API int b200_gemm_bf16(const void* A, int lda, const void* B, int ldb, const float* bias, void* D, int ldd,
                   int out_dtype, void* D2, int ldd2, int M, int N, int K, int relu, b200_stream_t stream);

/* b200_gemm_bf16 with two backward-pass epilogue options:
 *   accumulate != 0 : D (fp32) += result  — several producers add into one gradient buffer (e.g. dL/dx);
 *   mask (bf16, same shape as D, row stride ldmask) : result is zeroed where mask <= 0 — the ReLU backward of the
 *   layer whose forward activation `mask` is (attentive_modules.py:166-171, :72). */
B200_API int b200_gemm_bf16_ex(const void* A, int lda, const void* B, int ldb, const float* bias, void* D, int ldd,
                      int out_dtype, void* D2, int ldd2, int M, int N, int K, int relu, int accumulate,
                      const void* mask, int ldmask, b200_stream_t stream);

/* ---------------------------------------------------------------------------------------------------
 * P2 + the wide products of A1..A5 and their gradients: CTA-pair tcgen05 GEMM / implicit-GEMM convolution.
 *   D[M,N] = epi( [A | A2][M, K (+K2)] * B[N, K (+K2)]^T )      bf16 operands, fp32 accumulation
 * Replaces, with FrozenBN folded into the weights, the detectron2 BottleneckBlock convolutions the reference builds at
 * defrcn/modeling/roi_heads/roi_heads.py:313-337 and runs at :339-344 (1x1 convolutions = GEMMs over the NHWC pixels;
 * the 3x3 convolution = implicit GEMM whose A operand is fetched tap by tap by TMA with out-of-bounds zero fill as the
 * padding), their data gradients (autograd of the same), the `mean(dim=[2,3])` of roi_heads.py:1109 (rowmean_out), and
 * the nn.Linear products of attentive_modules.py:123-126,166-175,71-75 / their dX = dY W and dW = dY^T X without
 * transposed copies (a_mn / b_mn: the operand is stored M- / N-contiguous, i.e. transposed, in memory).
 *
 *   A         bf16.  a_mn = 0: [M][lda] (K contiguous);  a_mn = 1: [K][lda] (M contiguous);
 *             conv_c > 0: NHWC activation (M/16, 4, 4, conv_c) contiguous, K = 9 * conv_c ordered (tap = ky*3+kx, channel),
 *             stride 1, padding 1 (lda unused)
 *   A2, K2    optional second K segment (K-major, [M][lda2]); requires K % 64 == 0
 *   B         bf16.  b_mn = 0: [N][ldb] (K contiguous, K + K2 columns);  b_mn = 1: [K][ldb] (N contiguous)
 *   epilogue  v = acc + bias[n] (fp32, optional) + residual[m][n] (bf16, optional);  relu != 0: v = max(v, 0);
 *             mask_bits (packed words [m][n/32], layout below) or mask_act (bf16 activation): v = 0 where the bit is clear /
 *             the activation is <= 0  (ReLU backward)
 *   outputs   out_bf16 / out2_bf16 [M][ld] (TMA stores), out_f32 [M][ld_out_f32] (accumulate != 0: +=),
 *             bits_out: packed (v > 0) [M][ld_bits_out] words (needs N % 32 == 0); word w covers columns 32w .. 32w+31 with
 *             bit j <-> column 32w + 2j and bit 16 + j <-> column 32w + 2j + 1 (the order in which the epilogue's packed
 *             bf16x2 compare produces them); mask_bits uses the same layout,
 *             rowmean_out [M/16][ld_rowmean] fp32: mean of v over each group of 16 consecutive rows (the 4x4 pixels of a ROI)
 *   row-wise  (an epilogue lane owns one output row, so these cost no cross-lane traffic; they exclude the gates, bits_out,
 *   epilogues  rowmean_out, accumulate and split-K):
 *             rowsumsq_out [M][ld_rowsumsq] fp32: entry j of a row = sum of v^2 over columns 64j .. 64j+63 — the squared
 *             norm of an output row as a by-product (ceil(N/64) entries);
 *             row_scale_sumsq [M][ld] + row_scale_parts + row_scale_eps: acc is multiplied by
 *             1 / max(sqrt(sum of the row's first row_scale_parts entries), eps) before the bias — with the previous
 *             product's rowsumsq_out this is the cosine logit x.t / (|x| temperature) of my_module.py:449-469 without a
 *             normalisation pass (the temperature and 1/|t| are folded into B's rows);
 *             softmax != 0 (N <= 128): out_bf16 / out_f32 = softmax over the row of acc + bias — the attention
 *             probabilities of attentive_modules.py:45-55 as the epilogue of the (folded) score product;
 *             gate != 0: the residual operand is x and the outputs are out_bf16 = acc * x, out2_bf16 = x - acc — the gate
 *             operands of attentive_modules.py:166,170 as the epilogue of the probabilities x values product
 *   tile_n    0 = choose, 128 or 256;  max_clusters: 0 = one CTA pair per SM pair (tests lower it);
 *             epilogue_variant: 0 = the smallest compiled epilogue that covers the requested features, 1 = the generic one
 *             (same results; tests compare them)
 *   no_pdl    0: the kernel is launched with programmatic stream serialisation — its CTAs may become resident and set themselves
 *             up (barriers, TMEM) while the previous kernel of the stream drains, and wait (griddepcontrol.wait) for that
 *             kernel's completion before touching global memory; results are identical either way
 *   split_k   > 1: the K blocks of every output tile are shared by split_k CTA pairs (products with few output tiles and a
 *             long K: the (K+2)-row text operands, cls_score / bbox_pred and their weight gradients).  Each slice leaves its
 *             fp32 partial tile in splitk_workspace (b200_gemm2_splitk_workspace_bytes; its first 64 KiB hold the arrival
 *             counters, which must be zero on entry and are left zero, so one zero-initialised workspace serves any
 *             sequence of launches on one stream); the slice arriving last sums the partials in slice order — bitwise
 *             reproducible, no floating-point atomics — and runs the epilogue.  Needs an explicit tile_n, no residual.
 * Alignment: bf16 tensors 16-byte aligned, their leading dimensions multiples of 8 elements.  bf16 outputs leave through TMA
 * stores, which are 16-byte granular: when N is not a multiple of 8, the up to 7 elements that complete the last 16-byte unit
 * of a row (inside the row pitch, never in the next row) are written as well — with the epilogue's value for a column >= N
 * (bias-free zero accumulator: 0 after ReLU / softmax / gate).  fp32 outputs are exact-width.
 * ------------------------------------------------------------------------------------------------- */
typedef struct b200_gemm2_desc {
  const void* A; int lda;
  const void* A2; int lda2; int K2;
  const void* B; int ldb;
  int M, N, K;
  int a_mn, b_mn, conv_c;
  const float* bias;
  const void* residual; int ld_res;
  int relu;
  const void* mask_act; int ld_mask;
  const void* mask_bits; int ld_mask_bits;
  void* out_bf16; int ld_out;
  void* out2_bf16; int ld_out2;
  float* out_f32; int ld_out_f32; int accumulate;
  void* bits_out; int ld_bits_out;
  float* rowmean_out; int ld_rowmean;
  float* rowsumsq_out; int ld_rowsumsq;                 /* [M][ld]: sum of squares of the row's outputs per 64-column chunk */
  const float* row_scale_sumsq; int ld_row_scale_sumsq, row_scale_parts; float row_scale_eps;
                                                        /* acc *= 1 / max(sqrt(sum of the row's parts), eps), before the bias */
  int softmax;                                          /* out = softmax over the row (N <= 128) of acc + bias: bf16 and / or fp32 */
  int gate;                                             /* residual = x: out = acc * x, out2 = x - acc */
  int tile_n, max_clusters;
  int epilogue_variant;
  int no_pdl;                                           /* != 0: plain stream serialisation instead of a programmatic dependent launch */
  int split_k; void* splitk_workspace; size_t splitk_workspace_bytes;
} b200_gemm2_desc;
B200_API int b200_gemm2(const b200_gemm2_desc* desc, b200_stream_t stream);
B200_API size_t b200_gemm2_splitk_workspace_bytes(int M, int N, int tile_n, int split_k);

/* ---------------------------------------------------------------------------------------------------
 * Fine-tuning direction (BASELINE configs[1]): what autograd does for the reference's torch modules.
 * All reductions are fixed-order (deterministic); gradients feed the GEMM as bf16, accumulate in fp32.
 * ------------------------------------------------------------------------------------------------- */
/* dst[c][r] = bf16(src[r][c]) — operand preparation for the weight-gradient GEMMs dW = dY^T X (both operands
 * of b200_gemm_bf16 are K-contiguous, so dY^T and X^T are materialised) and for dX = dY W (W^T). */
B200_API int b200_transpose_bf16(const void* src, int src_dtype, int ld_src, void* dst, int ld_dst, int rows, int cols,
                        b200_stream_t stream);
/* out[c] (+)= sum_r src[r][c] — bias gradients (two-pass, ordered) */
B200_API size_t b200_colsum_workspace_bytes(int cols);
B200_API int b200_colsum(const void* src, int src_dtype, int ld, int rows, int cols, float* out, int accumulate,
                void* workspace, size_t workspace_bytes, b200_stream_t stream);
/* classifier dropout (fast_rcnn.py:412-414): y = bf16(keep ? x / (1-p) : 0), keep = hash(seed, index) >= p*2^32.
 * The mask is never stored: the backward kernel re-evaluates the hash.  seed_salt (device pointer, may be NULL) is
 * added to `seed` on the device: a step counter that lives in device memory lets a captured CUDA graph draw a new
 * mask on every replay. */
B200_API int b200_dropout_fwd(const float* x, void* y_bf16, size_t n, float p, unsigned long long seed,
                     const unsigned long long* seed_salt, b200_stream_t stream);
/* zd = bf16(dropout(relu?(LayerNorm(y + y2) * gamma + beta))): b200_residual_layernorm followed by b200_dropout_fwd in one pass,
 * bit for bit (same statistics, same counter-based keep decision hash(seed + *seed_salt, row * d + col)); the fp32 intermediate
 * is never written.  attentive_modules.py:73-74,285 + fast_rcnn.py:412-414. */
B200_API int b200_residual_layernorm_dropout(const float* y, const float* y2, const float* gamma, const float* beta, float eps,
                                    int relu, float p, unsigned long long seed, const unsigned long long* seed_salt,
                                    void* out_bf16, int R, int d, b200_stream_t stream);
/* backward of zd = dropout(relu(LayerNorm(y + y2))) (attentive_modules.py:73-74,285): du = dL/d(y + y2) as fp32
 * and/or bf16, dgamma / dbeta (may be NULL). */
B200_API size_t b200_layernorm_bwd_workspace_bytes(int R, int d);
B200_API int b200_layernorm_relu_dropout_bwd(const void* dzd_bf16, const float* y, const float* y2, const float* gamma,
                                    const float* beta, float eps, float p, unsigned long long seed,
                                    const unsigned long long* seed_salt, float* du_f32, void* du_bf16, float* dgamma,
                                    float* dbeta, int R, int d, void* workspace, size_t workspace_bytes,
                                    b200_stream_t stream);
/* dgamma / dbeta of the LayerNorm backward above, computed from the row statistics that call left in `workspace` (same
 * buffer, same arguments): separate so that it can run off the critical path, on another stream. */
B200_API int b200_layernorm_param_grads(const void* dzd_bf16, const float* y, const float* y2, const float* gamma,
                               const float* beta, float p, unsigned long long seed, const unsigned long long* seed_salt,
                               float* dgamma, float* dbeta, int R, int d, void* workspace, size_t workspace_bytes,
                               b200_stream_t stream);
/* backward of b200_text_attention (attentive_modules.py:45-55,166,170): given dP1, dP2 (bf16, row stride ldp) and
 * an optional external gradient on the attention probabilities (loss_attentive, roi_heads.py:1079-1081) computes
 *   dx (R,d) fp32 (+= when accumulate_dx), dO (R,d) bf16 (for dVp = attn^T dO through the GEMM),
 *   dS (R,ldds) bf16, zero padded beyond L (the scores' gradient: feeds dx += dS Kq and dKq = dS^T x). */
B200_API int b200_text_attention_bwd(const void* dp1, const void* dp2, int ldp, const float* x, const float* attn,
                            const float* vp, const float* dattn_ext, float* dx, int accumulate_dx, void* d_o, void* ds,
                            int ldds, int R, int d, int L, b200_stream_t stream);
/* L1: out3 = {loss_cls, loss_box_reg, loss_attentive} (fast_rcnn.py:222-304, roi_heads.py:1079-1081); attn may be
 * NULL (no attentive loss).  gt_classes int64 in [0, K] (K = background); proposals / gt_boxes (R,4).
 * acc_stats5 (optional): the counts FastRCNNOutputs._log_accuracy reads back (fast_rcnn.py:191-220) as floats:
 * {#(argmax == gt), #fg, #(fg and argmax == gt), #(fg and argmax == K), R}. */
B200_API int b200_head_losses(const float* logits, const float* deltas, const float* attn, const int64_t* gt_classes,
                     const float* proposals, const float* gt_boxes, int R, int K, int L, int cls_agnostic, float wx,
                     float wy, float ww, float wh, float smooth_l1_beta, float* out3, float* acc_stats5, b200_stream_t stream);
/* gradients of the three losses scaled by grad_scale3[0..2] (device): dlogits (R,ldl) bf16, ddeltas (R,ldd) bf16
 * (both zero padded to their row stride), dattn (R,L) fp32 (may be NULL). */
B200_API int b200_head_losses_bwd(const float* logits, const float* deltas, const float* attn, const int64_t* gt_classes,
                         const float* proposals, const float* gt_boxes, const float* grad_scale3, int R, int K, int L,
                         int cls_agnostic, float wx, float wy, float ww, float wh, float smooth_l1_beta,
                         void* dlogits_bf16, int ldl, void* ddeltas_bf16, int ldd, float* dattn, b200_stream_t stream);
/* Distillation loss of the student head — my_module.py:409-437 loss_fn_kd_only as called at roi_heads.py:760
 * (BASELINE configs[3]): out1[0] = alpha T^2 / R sum_r w_r KL(softmax(teacher_r / T) || softmax(student_r / T)),
 * w_r = 1.5 where gt_r == bg_label.  logits (R, C1) fp32.  b200_kd_loss_bwd ADDS grad_scale1[0] (device) times its
 * gradient to the bf16 student-logit gradient (R, ldl) that b200_head_losses_bwd left. */
B200_API int b200_kd_loss(const float* student_logits, const float* teacher_logits, const int64_t* gt_classes, int R, int C1,
                 int bg_label, float temperature, float alpha, float* out1, b200_stream_t stream);
B200_API int b200_kd_loss_bwd(const float* student_logits, const float* teacher_logits, const int64_t* gt_classes,
                     const float* grad_scale1, int R, int C1, int bg_label, float temperature, float alpha,
                     void* dlogits_bf16, int ldl, b200_stream_t stream);
/* T1 + text half of A1/A2 (attentive_modules.py:274-277, :125-135; Kq = Kp Wq / sqrt(d)): fp32 contractions with at
 * most 32 rows on one side, forward and backward (tensor-core tiles would be > 80 % padding; cuBLAS takes 30-50 us per
 * call here).  mode 0 "NT": out[m][n] = act(sum_k A[m][k] B[n][k] + bias[n]);  mode 1 "NN": out[m][k] = scale *
 * sum_n A'[m][n] B[n][k];  mode 2 "TN": out[n][k] (+)= sum_m A'[m][n] B[m][k], out_bias[n] (+)= sum_m A'[m][n].
 * A' = A zeroed where relu_ref <= 0 (ReLU backward; relu_ref may be NULL).  M <= 32 per call. */
B200_API size_t b200_skinny_gemm_workspace_bytes(int M, int cols);   /* cols = N (NT) or K (NN); TN needs none */
B200_API int b200_skinny_gemm(int mode, const float* A, int lda, const float* relu_ref, int ldref, const float* B, int ldb,
                     const float* bias, int relu, float scale, float* out, int ldo, float* out_bias, int M, int N, int K,
                     int accumulate, void* workspace, size_t workspace_bytes, b200_stream_t stream);
/* torch.optim.SGD step over one flat fp32 buffer: g += wd*p; m = mu*m + g; p -= lr*m (defrcn/solver/build.py).
 * bf16_shadow (optional, n elements): receives bf16(p) — the operand copy the tensor-core GEMMs of the next step read. */
B200_API int b200_sgd_momentum(float* params, const float* grads, float* momentum_buf, size_t n, float lr, float momentum,
                      float weight_decay, void* bf16_shadow, b200_stream_t stream);

/* A3 + A4 prologue: S = Q Kp^T / sqrt(d), softmax over the K+2 keys, O = attn Vp, then the two gate
 * operands P1 = O*x and P2 = x - O written as bf16 (attentive_modules.py:45-55,166,170).
 *   q (R,d) bf16 with kp (L,d) fp32, OR q == NULL and scores_in (R,L) fp32 = the already scaled scores S
 *   (the folded form S = x (Kp Wq)^T / sqrt(d), one skinny tensor-core GEMM instead of the d x d query GEMM);
 *   x (R,d) fp32 or bf16 (x_dtype); vp (L,d) fp32 (L = K+2, dummy key last, its value row 0);
 *   attn_out (R,L) fp32; p1,p2 (R, ldp) bf16. */
B200_API int b200_text_attention(const void* q, const float* scores_in, const void* x, int x_dtype, const float* kp,
                        const float* vp, float* attn_out, void* p1, void* p2, int ldp, int R, int d, int L,
                        b200_stream_t stream);

/* A5 tail + A6: out = relu?(LayerNorm(y + y2) * gamma + beta); out_f32 and/or out_bf16 may be NULL
 * (attentive_modules.py:73-74,285). */
B200_API int b200_residual_layernorm(const float* y, const float* y2, const float* gamma, const float* beta, float eps,
                            int relu, float* out_f32, void* out_bf16, int R, int d, b200_stream_t stream);

/* ---------------------------------------------------------------------------------------------------
 * Elementwise passes around the frozen res5 convolutions (roi_heads.py:313-344, spatial mean at :1109); all tensors
 * bf16 channels-last (R, HW, C) unless noted, C % 8 == 0, 16-byte aligned.
 *   b200_spatial_mean        pooled[r,c] (fp32, row stride ld_pooled) = mean over the HW pixels  — x.mean(dim=[2,3])
 *   b200_mean_bwd_relu_mask  g[r,p,c] = out[r,p,c] > 0 ? bf16(gpooled[r,c] / HW) : 0  — backward of the mean fused with
 *                            the ReLU backward of the last bottleneck's output `out`
 *   b200_add_relu_mask       y[i] = ref[i] > 0 ? bf16(a[i] + b[i]) : 0 over n elements — residual fan-in of a bottleneck
 *                            fused with the ReLU backward of the previous block's output `ref`; b and/or ref may be NULL
 * ------------------------------------------------------------------------------------------------- */
B200_API int b200_spatial_mean(const void* x_bf16, float* pooled, int ld_pooled, int R, int HW, int C, b200_stream_t stream);
B200_API int b200_mean_bwd_relu_mask(const float* gpooled, int ld_g, const void* out_bf16, void* g_bf16, int R, int HW, int C,
                            b200_stream_t stream);
B200_API int b200_add_relu_mask(const void* a_bf16, const void* b_bf16, const void* ref_bf16, void* y_bf16, size_t n,
                       b200_stream_t stream);

/* 1-bit ReLU masks for the passes above: relu_bits holds one byte per 8 consecutive elements of the activation (bit k
 * <-> element 8 i + k passes the ReLU backward), 1/16 of the activation's bytes.  b200_spatial_mean_bits = b200_spatial_mean
 * that also writes the mask of the tensor it averages; b200_pack_relu_bits writes the mask of any bf16 tensor (n elements,
 * n % 8 == 0); b200_mean_bwd_relu_bits / b200_add_relu_bits = the _mask entry points reading the mask instead of the
 * activation (add_relu_bits with a == the gradient, b == NULL is the in-place-capable ReLU backward y = a where kept).
 * b200_mean_bwd_relu_bits, bit_layout 1: the mask words b200_gemm2 writes (bits_out; C % 32 == 0): per 32 consecutive
 * channels one 32-bit word, bit j <-> channel 2j, bit 16 + j <-> channel 2j + 1. */
B200_API int b200_spatial_mean_bits(const void* x_bf16, float* pooled, int ld_pooled, void* relu_bits, int R, int HW, int C,
                           b200_stream_t stream);
B200_API int b200_pack_relu_bits(const void* x_bf16, void* relu_bits, size_t n, b200_stream_t stream);
B200_API int b200_mean_bwd_relu_bits(const float* gpooled, int ld_g, const void* relu_bits, void* g_bf16, int R, int HW, int C, int bit_layout,
                            b200_stream_t stream);
B200_API int b200_add_relu_bits(const void* a_bf16, const void* b_bf16, const void* relu_bits, void* y_bf16, size_t n,
                       b200_stream_t stream);

/* dst[r][:] = bf16(scale * src[r][:] / max(||src[r]||_2, eps)) — row normalisation for the optional cosine + temperature
 * form of the prototype logits (the reference's helper: my_module.py:461-469 `sim_matrix`, :449-458 `bsim_matrix`);
 * src fp32 or bf16 (src_dtype), leading dimensions in elements. */
B200_API int b200_l2_normalize_rows(const void* src, int src_dtype, int ld_src, void* dst_bf16, int ld_dst, int rows, int cols,
                           float eps, float scale, b200_stream_t stream);

/* ---------------------------------------------------------------------------------------------------
 * S1  ROIHeads.label_and_sample_proposals — defrcn/modeling/roi_heads/roi_heads.py:118-250
 *     (detectron2 pairwise_iou + Matcher([0.5],[0,1]) + subsample_labels), one launch for the whole batch.
 *   proposals (sum P_i, 4) / gt_boxes (sum M_i, 4) fp32 xyxy grouped by image with int32 offsets (num_images + 1);
 *   P_i <= 4096, M_i <= 256.  Per image: best ground-truth match per proposal (first maximum; foreground iff
 *   IoU >= iou_thresh; optional outputs matched_idx / matched_label over all proposals, bit-exact with the reference),
 *   then min(max_positive, #fg) foreground + min(batch - that, #bg) background proposals chosen uniformly at random
 *   (counter-based hash of (seed, image, proposal): the reference's distribution, not torch's RNG stream),
 *   foreground rows first.  Outputs have batch_per_image rows per image: sampled_idx (index within the image, -1 =
 *   padding), the sampled boxes, their class (num_classes = background, -1 = padding), the matched ground-truth box;
 *   counts (num_images, 2) = (#foreground rows, #valid rows).  append_gt != 0: the image's ground-truth boxes are
 *   candidates too, after its proposals (PROPOSAL_APPEND_GT, roi_heads.py:185-186; candidate index P_i + j = gt box j;
 *   matched_idx / matched_label then cover P_i + M_i entries per image), without a concatenated copy.  pad_background
 *   != 0: padding rows get class num_classes instead of -1 (fixed-shape consumers).  seed_salt (optional device scalar) is added to seed
 *   on the device: a step counter that survives CUDA-graph replay.  An image whose offsets exceed the limits gets no
 *   rows (counts 0) instead of being processed.
 * ------------------------------------------------------------------------------------------------- */
B200_API int b200_label_sample_proposals(const float* proposals, const int32_t* prop_offsets, const float* gt_boxes,
                                const int64_t* gt_classes, const int32_t* gt_offsets, int num_images,
                                int max_props_per_image, int max_gt_per_image, int num_classes, float iou_thresh,
                                int batch_per_image, int max_positive, unsigned long long seed, const int64_t* seed_salt, int append_gt, int pad_background, int32_t* matched_idx,
                                int32_t* matched_label, int32_t* sampled_idx, float* out_proposals, int64_t* out_classes,
                                float* out_gt_boxes, int32_t* counts, b200_stream_t stream);

/* ---------------------------------------------------------------------------------------------------
 * SURVEY 8f-3  RPN proposal selection == detectron2 0.3 find_top_rpn_proposals, vendored at
 *     defrcn/modeling/proposal_generator/proposal_utils.py:13-118 — the stage that feeds the ROI head.
 *   proposals (N, A, 4) fp32 xyxy (decoded anchors, levels concatenated along A), logits (N, A) fp32;
 *   level_offsets (L + 1) int32 DEVICE array, level l = anchors [level_offsets[l], level_offsets[l+1]).
 *   Per image and level: the min(pre_nms_topk, A_l) highest logits in descending order (ties: lower anchor
 *   index first; NaN sorts first like torch.sort); then per image: drop non-finite candidates (their number
 *   is written to n_invalid (N) — the reference raises FloatingPointError in training when it is non-zero),
 *   clip to image_hw (N,2) = (h, w), drop boxes whose width or height is not > min_box_size,
 *   batched_nms by level (coordinate-offset trick below 40000 candidates) at nms_thresh, keep the first
 *   post_nms_topk.  out_boxes (N, post_nms_topk, 4), out_logits (N, post_nms_topk) zero padded, out_count (N).
 *   cap_per_image = sum_l min(pre_nms_topk, A_l) (the caller knows the level sizes); pre_nms_topk <= 16384.  A smaller
 *   cap_per_image never overruns the workspace: candidates beyond it are dropped and n_invalid[image] = -(number dropped).
 *   Keep indices / counts are bit-exact with the reference's CPU path.  No host synchronisation.
 * ------------------------------------------------------------------------------------------------- */
B200_API size_t b200_rpn_select_workspace_bytes(int N, int cap_per_image, int L, int post_nms_topk);
B200_API int b200_rpn_select_proposals(const float* proposals, const float* logits, const int32_t* level_offsets,
                              const float* image_hw, int N, int A, int L, int pre_nms_topk, int post_nms_topk,
                              int cap_per_image, float nms_thresh, float min_box_size, float* out_boxes,
                              float* out_logits, int32_t* out_count, int32_t* n_invalid, void* workspace,
                              size_t workspace_bytes, b200_stream_t stream);

/* ---------------------------------------------------------------------------------------------------
 * SURVEY 8f-4  detectron2 0.3 detector_postprocess (call site defrcn/modeling/meta_arch/rcnn.py:69-73) on the
 *     padded detection tensors of b200_gather_detections, in place: boxes (N,max_keep,4) scaled by
 *     scale_xy (N,2) = (out_w / w, out_h / h) rounded to fp32, clipped to out_hw (N,2) = (out_h, out_w),
 *     empty boxes dropped (order kept; scores / classes / roi_inds rows move with their box; classes and
 *     roi_inds may be NULL); counts (N) updated.
 * ------------------------------------------------------------------------------------------------- */
B200_API int b200_detector_postprocess(float* boxes, float* scores, int64_t* classes, int64_t* roi_inds,
                              int32_t* counts, const float* scale_xy, const float* out_hw, int N, int max_keep,
                              b200_stream_t stream);

/* A7  teacher attention (attentive_modules.py:380-401 `one_hot(label) @ table`): dst[r][:] = bf16(act(table[idx[r]][:])),
 * table (num_rows_table, cols) fp32, idx (rows) int64 clamped to the table, dst bf16 with row stride ld_dst. */
B200_API int b200_gather_rows_bf16(const float* table, int ld_table, int num_rows_table, const int64_t* idx, void* dst,
                          int ld_dst, int rows, int cols, int relu, b200_stream_t stream);

/* A7  teacher attention: mean (C,d) = per-class mean of the rows of x (R,d) fp32 (row stride ldx) grouped by labels (R)
 * int64 in [0, C), counts (C) fp32 = rows per class (classes without rows: mean 0).  Fixed summation order. */
B200_API size_t b200_class_mean_rows_workspace_bytes(int R, int d, int C);
B200_API int b200_class_mean_rows(const float* x, int ldx, const int64_t* labels, int R, int d, int C, float* mean,
                         float* counts, void* workspace, size_t workspace_bytes, b200_stream_t stream);

/* fp32 -> bf16 cast with row stride (builds the [o1|o2|x] concat buffer of attentive_modules.py:172-174
 * in place, without a torch.cat) */
B200_API int b200_cast_bf16(const float* src, int ld_src, void* dst, int ld_dst, int rows, int cols,
                   b200_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* B200ROI_H_ */
