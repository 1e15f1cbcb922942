// P1b: ROIAlign backward, atomic-free and deterministic (the reference's torchvision kernel scatters with
// atomicAdd — defrcn fine-tuning reaches it through autograd of roi_heads.py:340).
//
// Gather formulation.  ROIAlign with sample averaging is separable:
//   out[r,ph,pw,c] = 1/count_r * sum_{y,x} A_r[ph][y] * B_r[pw][x] * feat[y,x,c]
// with A_r[ph][y] the summed vertical bilinear weights of bin ph's samples on row y (B_r likewise), so
//   grad_feat[n,y,x,c] = sum_{r in image n} 1/count_r * sum_{ph,pw} A_r[ph][y] * B_r[pw][x] * g[r,ph,pw,c].
// Kernel 1 builds the (tiny) per-ROI weight tables, one thread per (roi, bin) so each table row has a single
// writer.  Kernel 2 assigns one warp to a feature-map pixel and 4*32 consecutive channels; it walks the
// image's ROIs in index order (a segmented reduction over the ROIs covering the pixel), accumulates in
// registers and writes every grad_feat element exactly once — no zero-fill, no atomics, run-to-run identical.
#include "common.cuh"

namespace b200 {

int dispatch_affine(const void* x, const float* w, const float* b, float mult, void* y, int N, int C, int H, int W,
                    int in_dtype, int in_layout, int out_dtype, int out_layout, cudaStream_t st);

constexpr int kMaxP = 8;  // pooled size supported by the backward (the head uses 7)

// bf16 channels-last 7x7 path — 2: pixel-tile gather on the tensor cores (roi_align_bwd_tile.cu; C % 64 == 0, else 1),
// 1: per-pixel CSR gather (roi_align_bwd_slice.cu), 0: the table kernel below
int g_roi_bwd_impl = 2;
bool roi_bwd_tile_eligible(int C, int H, int W, int PH, int PW, int bin_step);
size_t roi_bwd_tile_workspace_bytes(int N, int H, int W, int R, int PH, int PW, int bin_step);
int launch_roi_bwd_tile_plan(const float* rois, const int32_t* roi_offsets, int N, int H, int W, int R, int PH, int PW,
                             int bin_step, float scale, int sr, int aligned, void* workspace, cudaStream_t st);
int launch_roi_bwd_tile_gather(const __nv_bfloat16* g, const void* workspace, __nv_bfloat16* grad_feat, int N, int C, int H,
                               int W, int R, int PH, int PW, int bin_step, cudaStream_t st);
bool roi_bwd_slice_eligible(int C, int H, int W, int PH, int PW, int bin_step);
size_t roi_bwd_slice_workspace_bytes(int N, int H, int W, int R, int PH, int PW, int bin_step);
int launch_roi_bwd_plan(const float* rois, const int32_t* roi_offsets, int N, int H, int W, int R, int PH, int PW,
                        int bin_step, float scale, int sr, int aligned, void* workspace, cudaStream_t st);
int launch_roi_bwd_gather(const __nv_bfloat16* g, const void* workspace, __nv_bfloat16* grad_feat, int N, int C, int H, int W,
                          int R, int PH, int PW, int bin_step, cudaStream_t st);
int launch_roi_bwd_slice(const __nv_bfloat16* g, const float* rois, const int32_t* roi_offsets, __nv_bfloat16* grad_feat,
                         int N, int C, int H, int W, int R, int PH, int PW, int bin_step, float scale, int sr, int aligned,
                         void* workspace, cudaStream_t st);

struct RoiExtent {
  int ylo, yhi, xlo, xhi;  // inclusive pixel ranges with non-zero weight (ylo > yhi: empty)
  float inv_count;
  int pad[3];
};

struct BwdTables {
  float* ta;        // [R][H][kMaxP]  A_r[ph][y] stored ph-minor
  float* tb;        // [R][W][kMaxP]
  uint16_t* ra;     // [R][H]  ph_lo | ph_hi << 8
  uint16_t* rb;     // [R][W]
  RoiExtent* ext;   // [R]
};

__device__ __forceinline__ float bwd_sample_coord(float start, int p, float bin, int i, int grid) {
  return __fadd_rn(__fadd_rn(start, __fmul_rn((float)p, bin)), __fdiv_rn(__fmul_rn((float)i + .5f, bin), (float)grid));
}

// grid = R, block = 64.  threads [0,PH) build A rows, [32,32+PW) build B rows.
__global__ void roi_bwd_tables_kernel(const float* __restrict__ rois, BwdTables t, int H, int W, int PH, int PW,
                                      float scale, int sampling_ratio, int aligned) {
  const int r = blockIdx.x;
  const float* roi = rois + 5 * (size_t)r;
  const float off = aligned ? 0.5f : 0.0f;
  const float sw = __fsub_rn(__fmul_rn(roi[1], scale), off), sh = __fsub_rn(__fmul_rn(roi[2], scale), off);
  const float ew = __fsub_rn(__fmul_rn(roi[3], scale), off), eh = __fsub_rn(__fmul_rn(roi[4], scale), off);
  float rw = __fsub_rn(ew, sw), rh = __fsub_rn(eh, sh);
  if (!aligned) { rw = fmaxf(rw, 1.f); rh = fmaxf(rh, 1.f); }
  const float bin_h = __fdiv_rn(rh, (float)PH), bin_w = __fdiv_rn(rw, (float)PW);
  int gh = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(__fdiv_rn(rh, (float)PH));
  int gw = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(__fdiv_rn(rw, (float)PW));
  gh = max(gh, 0); gw = max(gw, 0);

  __shared__ int s_lo[2], s_hi[2];
  if (threadIdx.x < 2) { s_lo[threadIdx.x] = 1 << 30; s_hi[threadIdx.x] = -1; }
  __syncthreads();
  const int axis = threadIdx.x >> 5;  // 0: rows (A), 1: cols (B)
  const int p = threadIdx.x & 31;
  const int P = axis ? PW : PH, size = axis ? W : H, g = axis ? gw : gh;
  const float start = axis ? sw : sh, bin = axis ? bin_w : bin_h;
  float* tab = axis ? t.tb + (size_t)r * W * kMaxP : t.ta + (size_t)r * H * kMaxP;
  if (p < P) {
    int lo_seen = 1 << 30, hi_seen = -1;
    for (int i = 0; i < g; ++i) {
      float coord = bwd_sample_coord(start, p, bin, i, g);
      if (coord < -1.0f || coord > (float)size) continue;
      if (coord <= 0.f) coord = 0.f;
      int lo = (int)coord, hi;
      if (lo >= size - 1) { hi = lo = size - 1; coord = (float)lo; } else hi = lo + 1;
      const float l = coord - (float)lo;
      tab[(size_t)lo * kMaxP + p] += 1.f - l;   // single writer per (r, axis, p): plain adds, fixed order
      tab[(size_t)hi * kMaxP + p] += l;
      lo_seen = min(lo_seen, lo); hi_seen = max(hi_seen, hi);
    }
    if (hi_seen >= 0) { atomicMin(&s_lo[axis], lo_seen); atomicMax(&s_hi[axis], hi_seen); }  // integer: order-free
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    RoiExtent e;
    e.ylo = s_lo[0]; e.yhi = s_hi[0]; e.xlo = s_lo[1]; e.xhi = s_hi[1];
    if (e.yhi < 0 || e.xhi < 0) { e.ylo = 1; e.yhi = 0; e.xlo = 1; e.xhi = 0; }
    e.inv_count = 1.0f / (float)max(gh * gw, 1);
    e.pad[0] = e.pad[1] = e.pad[2] = 0;
    t.ext[r] = e;
  }
  // per-pixel bin ranges
  const int lo = s_lo[axis], hi = s_hi[axis];
  uint16_t* rng = axis ? t.rb + (size_t)r * W : t.ra + (size_t)r * H;
  for (int i = lo + p; i <= hi; i += 32) {
    int plo = 255, phi = 0;
    for (int k = 0; k < P; ++k)
      if (tab[(size_t)i * kMaxP + k] != 0.f) { plo = min(plo, k); phi = max(phi, k); }
    rng[i] = (uint16_t)(plo | (phi << 8));   // plo=255 > phi=0 when the pixel has no weight
  }
}

constexpr int kBwdWarps = 8;

template <typename T> struct V4;
template <> struct V4<float> {
  static __device__ __forceinline__ float4 ld(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
  static __device__ __forceinline__ void st(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
};
template <> struct V4<__nv_bfloat16> {
  static __device__ __forceinline__ float4 ld(const __nv_bfloat16* p) {
    const uint2 t = __ldg(reinterpret_cast<const uint2*>(p));
    return make_float4(__uint_as_float(t.x << 16), __uint_as_float(t.x & 0xffff0000u), __uint_as_float(t.y << 16),
                       __uint_as_float(t.y & 0xffff0000u));
  }
  static __device__ __forceinline__ void st(__nv_bfloat16* p, float4 v) {
    const __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 t;
    t.x = *reinterpret_cast<const uint32_t*>(&a); t.y = *reinterpret_cast<const uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = t;
  }
};

// grid (H, N, ceil(C/128)); block kBwdWarps warps.  g is NHWC: [R][PH][PW][C]; grad_feat NHWC.
template <typename T>
__global__ void __launch_bounds__(kBwdWarps * 32)
roi_align_bwd_nhwc_kernel(const T* __restrict__ g, const int32_t* __restrict__ roi_batch_offsets, BwdTables t,
                          T* __restrict__ grad_feat, int C, int H, int W, int PH, int PW, int bin_step) {
  extern __shared__ int s_list[];  // ROIs of this image whose row extent covers y
  __shared__ int s_n;
  const int y = blockIdx.x, n = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.z * 128 + lane * 4;
  const int r0 = roi_batch_offsets[n], r1 = roi_batch_offsets[n + 1];
  // bin_step > 1: g holds only the bins (0, step, 2*step, ...) of each axis, densely: [R][PHO][PWO][C]
  const int PHO = (PH + bin_step - 1) / bin_step, PWO = (PW + bin_step - 1) / bin_step;
  // ordered compaction of the covering ROIs (warp 0, ballot prefix) keeps the summation order fixed
  if (warp == 0) {
    int cnt = 0;
    for (int rb = r0; rb < r1; rb += 32) {
      const int r = rb + lane;
      bool hit = false;
      if (r < r1) { const RoiExtent e = t.ext[r]; hit = e.ylo <= y && y <= e.yhi; }
      const unsigned m = __ballot_sync(0xffffffffu, hit);
      if (hit) s_list[cnt + __popc(m & ((1u << lane) - 1u))] = r;
      cnt += __popc(m);
    }
    if (lane == 0) s_n = cnt;
  }
  __syncthreads();
  const int nlist = s_n;
  if (c >= C) return;
  for (int x = warp; x < W; x += kBwdWarps) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int li = 0; li < nlist; ++li) {
      const int r = s_list[li];
      const RoiExtent e = t.ext[r];
      if (x < e.xlo || x > e.xhi) continue;
      const uint16_t ry = t.ra[(size_t)r * H + y], rx = t.rb[(size_t)r * W + x];
      const int ph_lo = ry & 255, ph_hi = ry >> 8, pw_lo = rx & 255, pw_hi = rx >> 8;
      const float* wa = t.ta + ((size_t)r * H + y) * kMaxP;
      const float* wb = t.tb + ((size_t)r * W + x) * kMaxP;
      for (int ph = (ph_lo + bin_step - 1) / bin_step * bin_step; ph <= ph_hi; ph += bin_step) {
        const float wy = wa[ph] * e.inv_count;
        if (wy == 0.f) continue;
        const T* grow = g + (((size_t)r * PHO + ph / bin_step) * PWO) * C + c;
        for (int pw = (pw_lo + bin_step - 1) / bin_step * bin_step; pw <= pw_hi; pw += bin_step) {
          const float w = wy * wb[pw];
          if (w == 0.f) continue;
          const float4 gv = V4<T>::ld(grow + (size_t)(pw / bin_step) * C);
          acc.x += w * gv.x; acc.y += w * gv.y; acc.z += w * gv.z; acc.w += w * gv.w;
        }
      }
    }
    V4<T>::st(grad_feat + (((size_t)n * H + y) * W + x) * C + c, acc);
  }
}

static size_t tables_bytes(int R, int H, int W) {
  return align_up((size_t)R * H * kMaxP * 4, 256) + align_up((size_t)R * W * kMaxP * 4, 256) +
         align_up((size_t)R * H * 2, 256) + align_up((size_t)R * W * 2, 256) + align_up((size_t)R * sizeof(RoiExtent), 256);
}

}  // namespace b200

using namespace b200;

extern "C" size_t b200_roi_align_bwd_workspace_bytes(int N, int C, int H, int W, int R, int pooled_h, int pooled_w,
                                                     int bin_step, int dtype, int grad_in_layout, int grad_out_layout) {
  const size_t e = dtype == B200_BF16 ? 2 : 4;
  size_t b = tables_bytes(R, H, W);
  bin_step = max(bin_step, 1);
  const int pho = ceil_div(pooled_h, bin_step), pwo = ceil_div(pooled_w, bin_step);
  if (grad_out_layout == B200_NCHW) b += align_up((size_t)R * C * pho * pwo * e, 256);
  if (grad_in_layout == B200_NCHW) b += align_up((size_t)N * C * H * W * e, 256);
  if (dtype == B200_BF16 && roi_bwd_slice_eligible(C, H, W, pooled_h, pooled_w, bin_step))
    b = max(b, roi_bwd_slice_workspace_bytes(N, H, W, R, pooled_h, pooled_w, bin_step));
  if (dtype == B200_BF16 && roi_bwd_tile_eligible(C, H, W, pooled_h, pooled_w, bin_step))
    b = max(b, roi_bwd_tile_workspace_bytes(N, H, W, R, pooled_h, pooled_w, bin_step));
  return b;
}

extern "C" int b200_roi_align_bwd(const void* grad_out, const float* rois, const int32_t* roi_batch_offsets,
                                  void* grad_feat, int N, int C, int H, int W, int R, int pooled_h, int pooled_w,
                                  int bin_step, float spatial_scale, int sampling_ratio, int aligned, int dtype, int grad_out_layout,
                                  int grad_in_layout, void* workspace, size_t workspace_bytes, b200_stream_t stream) {
  B200_CHECK_ARG(grad_feat && roi_batch_offsets && (R == 0 || (grad_out && rois)), "roi_align_bwd: null tensor");
  B200_CHECK_ARG(N > 0 && C > 0 && H > 0 && W > 0 && R >= 0, "roi_align_bwd: bad shape");
  B200_CHECK_ARG(bin_step >= 1 && bin_step <= 8, "roi_align_bwd: bin_step must be in [1, 8]");
  const int pho = ceil_div(pooled_h, bin_step), pwo = ceil_div(pooled_w, bin_step);
  if (pooled_h > kMaxP || pooled_w > kMaxP || C % 4 != 0) {
    set_error("roi_align_bwd: pooled size must be <= %d and C %% 4 == 0", kMaxP);
    return B200_ERR_UNSUPPORTED;
  }
  const size_t need = b200_roi_align_bwd_workspace_bytes(N, C, H, W, R, pooled_h, pooled_w, bin_step, dtype, grad_in_layout, grad_out_layout);
  if ((need && !workspace) || workspace_bytes < need) {
    set_error("roi_align_bwd: workspace too small (%zu < %zu)", workspace_bytes, need);
    return B200_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const bool cl16 = R > 0 && dtype == B200_BF16 && grad_out_layout == B200_NHWC && grad_in_layout == B200_NHWC &&
                    (((uintptr_t)grad_out | (uintptr_t)grad_feat) & 15) == 0;
  if (cl16 && g_roi_bwd_impl == 2 && roi_bwd_tile_eligible(C, H, W, pooled_h, pooled_w, bin_step)) {
    int rc = launch_roi_bwd_tile_plan(rois, roi_batch_offsets, N, H, W, R, pooled_h, pooled_w, bin_step, spatial_scale,
                                      sampling_ratio, aligned, workspace, st);
    if (rc != B200_OK) return rc;
    return launch_roi_bwd_tile_gather((const __nv_bfloat16*)grad_out, workspace, (__nv_bfloat16*)grad_feat, N, C, H, W, R,
                                      pooled_h, pooled_w, bin_step, st);
  }
  if (cl16 && g_roi_bwd_impl >= 1 && roi_bwd_slice_eligible(C, H, W, pooled_h, pooled_w, bin_step))
    return launch_roi_bwd_slice((const __nv_bfloat16*)grad_out, rois, roi_batch_offsets, (__nv_bfloat16*)grad_feat, N, C, H, W,
                                R, pooled_h, pooled_w, bin_step, spatial_scale, sampling_ratio, aligned, workspace, st);
  unsigned char* p = (unsigned char*)workspace;
  BwdTables t;
  t.ta = (float*)p;      p += align_up((size_t)R * H * kMaxP * 4, 256);
  t.tb = (float*)p;      p += align_up((size_t)R * W * kMaxP * 4, 256);
  t.ra = (uint16_t*)p;   p += align_up((size_t)R * H * 2, 256);
  t.rb = (uint16_t*)p;   p += align_up((size_t)R * W * 2, 256);
  t.ext = (RoiExtent*)p; p += align_up((size_t)R * sizeof(RoiExtent), 256);
  const size_t e = dtype == B200_BF16 ? 2 : 4;
  const void* g = grad_out;
  if (R > 0) {
    B200_CUDA_CALL(cudaMemsetAsync(t.ta, 0, (size_t)((unsigned char*)t.ra - (unsigned char*)t.ta), st));
    roi_bwd_tables_kernel<<<R, 64, 0, st>>>(rois, t, H, W, pooled_h, pooled_w, spatial_scale, sampling_ratio, aligned);
    B200_CUDA_LAUNCH_CHECK("roi_bwd_tables");
    if (grad_out_layout == B200_NCHW) {
      int rc = dispatch_affine(grad_out, nullptr, nullptr, 1.0f, p, R, C, pho, pwo, dtype, B200_NCHW, dtype, B200_NHWC, st);
      if (rc != B200_OK) return rc;
      g = p;
      p += align_up((size_t)R * C * pho * pwo * e, 256);
    }
  }
  void* gf = grad_in_layout == B200_NCHW ? (void*)p : grad_feat;
  // largest per-image ROI count is not known on the host: size the list for all R
  const size_t smem = (size_t)max(R, 1) * sizeof(int);
  dim3 grid(H, N, ceil_div(C, 128));
  if (dtype == B200_F32) {
    auto k = roi_align_bwd_nhwc_kernel<float>;
    if (smem > 40 * 1024) B200_CUDA_CALL(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, kBwdWarps * 32, smem, st>>>((const float*)g, roi_batch_offsets, t, (float*)gf, C, H, W, pooled_h, pooled_w, bin_step);
  } else {
    auto k = roi_align_bwd_nhwc_kernel<__nv_bfloat16>;
    if (smem > 40 * 1024) B200_CUDA_CALL(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, kBwdWarps * 32, smem, st>>>((const __nv_bfloat16*)g, roi_batch_offsets, t, (__nv_bfloat16*)gf, C, H, W, pooled_h, pooled_w, bin_step);
  }
  B200_CUDA_LAUNCH_CHECK("roi_align_bwd");
  if (grad_in_layout == B200_NCHW)
    return dispatch_affine(gf, nullptr, nullptr, 1.0f, grad_feat, N, C, H, W, dtype, B200_NHWC, dtype, B200_NCHW, st);
  return B200_OK;
}

// ---- split form of the bf16 channels-last path: the geometry plan depends only on the ROIs, so it can be built ahead
// of the backward pass (on a side stream during the forward) and consumed by one gather launch ------------------------
extern "C" size_t b200_roi_align_bwd_plan_bytes(int N, int C, int H, int W, int R, int pooled_h, int pooled_w, int bin_step) {
  bin_step = max(bin_step, 1);
  if (N <= 0 || H <= 0 || W <= 0 || R < 0 || !roi_bwd_slice_eligible(C, H, W, pooled_h, pooled_w, bin_step)) return 0;
  // one size for both list formats, so that a plan buffer stays valid whichever "roi_align_bwd_impl" is selected
  size_t b = roi_bwd_slice_workspace_bytes(N, H, W, R, pooled_h, pooled_w, bin_step);
  if (roi_bwd_tile_eligible(C, H, W, pooled_h, pooled_w, bin_step))
    b = max(b, roi_bwd_tile_workspace_bytes(N, H, W, R, pooled_h, pooled_w, bin_step));
  return b;
}

extern "C" int b200_roi_align_bwd_plan(const float* rois, const int32_t* roi_batch_offsets, int N, int C, int H, int W, int R,
                                       int pooled_h, int pooled_w, int bin_step, float spatial_scale, int sampling_ratio,
                                       int aligned, void* plan, size_t plan_bytes, b200_stream_t stream) {
  B200_CHECK_ARG(roi_batch_offsets && (R == 0 || rois), "roi_align_bwd_plan: null tensor");
  B200_CHECK_ARG(N > 0 && C > 0 && H > 0 && W > 0 && R >= 0 && bin_step >= 1 && bin_step <= 8, "roi_align_bwd_plan: bad shape");
  const size_t need = b200_roi_align_bwd_plan_bytes(N, C, H, W, R, pooled_h, pooled_w, bin_step);
  if (need == 0) {
    set_error("roi_align_bwd_plan: shape not covered by the planned path (7x7 pooling, C %% 8 == 0, map <= 256 x 256)");
    return B200_ERR_UNSUPPORTED;
  }
  if (!plan || plan_bytes < need) {
    set_error("roi_align_bwd_plan: plan buffer too small (%zu < %zu)", plan_bytes, need);
    return B200_ERR_WORKSPACE;
  }
  // the plan's format follows "roi_align_bwd_impl" at the time of the call; b200_roi_align_bwd_planned must see the same
  if (g_roi_bwd_impl == 2 && roi_bwd_tile_eligible(C, H, W, pooled_h, pooled_w, bin_step))
    return launch_roi_bwd_tile_plan(rois, roi_batch_offsets, N, H, W, R, pooled_h, pooled_w, bin_step, spatial_scale,
                                    sampling_ratio, aligned, plan, (cudaStream_t)stream);
  return launch_roi_bwd_plan(rois, roi_batch_offsets, N, H, W, R, pooled_h, pooled_w, bin_step, spatial_scale, sampling_ratio,
                             aligned, plan, (cudaStream_t)stream);
}

extern "C" int b200_roi_align_bwd_planned(const void* grad_out, const void* plan, size_t plan_bytes, void* grad_feat, int N, int C,
                                          int H, int W, int R, int pooled_h, int pooled_w, int bin_step, b200_stream_t stream) {
  B200_CHECK_ARG(grad_feat && plan && (R == 0 || grad_out), "roi_align_bwd_planned: null tensor");
  B200_CHECK_ARG((((uintptr_t)grad_out | (uintptr_t)grad_feat) & 15) == 0, "roi_align_bwd_planned: tensors must be 16-byte aligned");
  const size_t need = b200_roi_align_bwd_plan_bytes(N, C, H, W, R, pooled_h, pooled_w, bin_step);
  if (need == 0 || plan_bytes < need) {
    set_error("roi_align_bwd_planned: plan buffer does not match the shape (%zu < %zu)", plan_bytes, need);
    return B200_ERR_WORKSPACE;
  }
  if (g_roi_bwd_impl == 2 && roi_bwd_tile_eligible(C, H, W, pooled_h, pooled_w, bin_step))
    return launch_roi_bwd_tile_gather((const __nv_bfloat16*)grad_out, plan, (__nv_bfloat16*)grad_feat, N, C, H, W, R, pooled_h,
                                      pooled_w, bin_step, (cudaStream_t)stream);
  return launch_roi_bwd_gather((const __nv_bfloat16*)grad_out, plan, (__nv_bfloat16*)grad_feat, N, C, H, W, R, pooled_h,
                               pooled_w, bin_step, (cudaStream_t)stream);
}
