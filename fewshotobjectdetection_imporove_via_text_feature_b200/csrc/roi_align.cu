// P1 / Q1: ROIAlign forward (torchvision roi_align semantics: aligned flag, adaptive sampling grid).
// Reference call sites: defrcn/modeling/roi_heads/roi_heads.py:300-305,339-340 (7x7 @ 1/16 on res4) and
// defrcn/evaluation/calibration_layer.py:27,100 (1x1 @ 1/32, PCB).
//
// Data layout in HBM: the gather runs on an NHWC (channels-last) map so that one bilinear tap of one
// sample is a fully coalesced 16-byte-per-lane read of consecutive channels; an NCHW map is first
// re-laid-out into the workspace (one extra read+write of the map, ~7 % of the algorithmic bytes at
// R=512).  The output — 93 % of the algorithmic bytes — is written either NHWC (coalesced 16 B streaming
// stores straight from registers) or NCHW (transposed through shared memory so that the CTA's
// [channels x bins] slab leaves as one contiguous burst).
//
// Work decomposition: CTA = (roi, chunk of 32*VEC channels), 8 warps; each warp owns output bins
// b = warp, warp+8, ...; per-ROI, per-axis bin windows (first pixel, count, pre-summed separable weights) are
// built once per CTA in shared memory, so the inner loop is 1 vector load + VEC FMAs per distinct pixel.
#include "common.cuh"

namespace b200 {

int dispatch_affine(const void* x, const float* w, const float* b, float mult, void* y, int N, int C, int H, int W,
                    int in_dtype, int in_layout, int out_dtype, int out_layout, cudaStream_t st);

struct AxisTap {
  int lo, hi;      // element offsets along the axis (already multiplied by the axis stride)
  float wlo, whi;  // (1-frac), frac ; both 0 when the sample is out of range
};

// One-axis half of torchvision's bilinear_interpolate / pre_calc_for_bilinear_interpolate.
__device__ __forceinline__ AxisTap make_tap(float coord, int size, int stride) {
  AxisTap t;
  if (coord < -1.0f || coord > (float)size) {
    t.lo = t.hi = 0; t.wlo = t.whi = 0.f;
    return t;
  }
  if (coord <= 0.f) coord = 0.f;
  int lo = (int)coord, hi;
  if (lo >= size - 1) { hi = lo = size - 1; coord = (float)lo; } else hi = lo + 1;
  const float l = coord - (float)lo;
  t.lo = lo * stride; t.hi = hi * stride; t.wlo = 1.f - l; t.whi = l;
  return t;
}

struct RoiGeom {
  int batch, gh, gw;
  float start_h, start_w, bin_h, bin_w, count;
};

__device__ __forceinline__ RoiGeom roi_geom(const float* __restrict__ roi, float scale, int sampling_ratio,
                                            int aligned, int PH, int PW) {
  RoiGeom g;
  g.batch = (int)roi[0];
  const float off = aligned ? 0.5f : 0.0f;
  // no FMA contraction: coordinates must round exactly as the CPU reference's
  g.start_w = __fsub_rn(__fmul_rn(roi[1], scale), off);
  g.start_h = __fsub_rn(__fmul_rn(roi[2], scale), off);
  const float end_w = __fsub_rn(__fmul_rn(roi[3], scale), off);
  const float end_h = __fsub_rn(__fmul_rn(roi[4], scale), off);
  float rw = __fsub_rn(end_w, g.start_w), rh = __fsub_rn(end_h, g.start_h);
  if (!aligned) { rw = fmaxf(rw, 1.f); rh = fmaxf(rh, 1.f); }
  g.bin_h = __fdiv_rn(rh, (float)PH);
  g.bin_w = __fdiv_rn(rw, (float)PW);
  g.gh = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(__fdiv_rn(rh, (float)PH));
  g.gw = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(__fdiv_rn(rw, (float)PW));
  g.gh = max(g.gh, 0); g.gw = max(g.gw, 0);
  g.count = (float)max(g.gh * g.gw, 1);
  return g;
}

__device__ __forceinline__ float sample_coord(float start, int p, float bin, int i, int grid) {
  // roi_start + p*bin + (i + .5f) * bin / grid   (left-to-right, separate roundings)
  return __fadd_rn(__fadd_rn(start, __fmul_rn((float)p, bin)),
                   __fdiv_rn(__fmul_rn((float)i + .5f, bin), (float)grid));
}

template <typename T, int VEC> struct Vec;
template <> struct Vec<float, 4> {
  float v[4];
  __device__ __forceinline__ void load(const float* p) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  __device__ static __forceinline__ void store_stream(float* p, const float* a) {
    st_stream_f4(reinterpret_cast<float4*>(p), make_float4(a[0], a[1], a[2], a[3]));
  }
};
template <> struct Vec<__nv_bfloat16, 8> {
  float v[8];
  __device__ __forceinline__ void load(const __nv_bfloat16* p) {
    const uint4 t = __ldg(reinterpret_cast<const uint4*>(p));
    const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[2 * i] = __uint_as_float(w[i] << 16);
      v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
  __device__ static __forceinline__ void store_stream(__nv_bfloat16* p, const float* a) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const __nv_bfloat162 h = __floats2bfloat162_rn(a[2 * i], a[2 * i + 1]);
      w[i] = *reinterpret_cast<const uint32_t*>(&h);
    }
    asm volatile("st.global.cs.v4.b32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3])
                 : "memory");
  }
};
template <typename T> __device__ __forceinline__ T from_float(float v);
template <> __device__ __forceinline__ float from_float<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_float<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

constexpr int kRoiWarps = 8;
constexpr int kMaxTaps = 192;  // per-axis weight slots P*(g+1) held in smem; larger ROIs take the per-sample path

// Per-bin pixel windows.  ROIAlign's sample average is separable: within bin (ph,pw) the gh x gw bilinear samples
// touch at most (gh+1) x (gw+1) distinct pixels, and pixel (y,x) carries weight a[ph][y] * b[pw][x] where a / b are
// the per-axis sums of the samples' (1-frac, frac) weights.  The tables below hold, per axis and per bin, the
// first pixel, the pixel count and the summed weights, so the inner loop visits every distinct pixel ONCE
// (1 vector load + VEC FMAs) instead of 4 taps per sample — ~2-3x fewer instructions on the large ROIs that
// dominate the work (ncu: the per-sample form was issue-bound, 86 % issue-active, DRAM 7 %).
struct AxisBins {
  int start[16];   // first pixel of the bin's window, in elements (already multiplied by the axis stride)
  int count[16];   // pixels in the window (0: every sample of the bin fell outside the map)
  float w[kMaxTaps];
};

__device__ __forceinline__ void build_axis_bins(AxisBins& t, int p, int P, int g, float start, float bin, int size,
                                                int stride) {
  // one thread per bin; weights were zeroed by the caller
  int first = -1, cnt = 0;
  float* w = t.w + p * (g + 1);
  for (int i = 0; i < g; ++i) {
    float coord = sample_coord(start, p, bin, i, g);
    if (coord < -1.0f || coord > (float)size) continue;
    if (coord <= 0.f) coord = 0.f;
    int lo = (int)coord, hi;
    if (lo >= size - 1) { hi = lo = size - 1; coord = (float)lo; } else hi = lo + 1;
    const float l = coord - (float)lo;
    if (first < 0) first = lo;
    w[lo - first] += 1.f - l;
    w[hi - first] += l;
    cnt = max(cnt, hi - first + 1);
  }
  t.start[p] = max(first, 0) * stride;
  t.count[p] = cnt;
  (void)P;
}

// OUT_MODE 0: NHWC direct, 1: NCHW through smem (dynamic smem: bins*(CH+1) floats), 2: NCHW direct scatter
template <typename T, int VEC, int OUT_MODE>
__global__ void __launch_bounds__(kRoiWarps * 32)
roi_align_fwd_nhwc_kernel(const T* __restrict__ feat, const float* __restrict__ rois, T* __restrict__ out, int C,
                          int H, int W, int PH, int PW, float scale, int sampling_ratio, int aligned) {
  constexpr int CH = 32 * VEC;
  extern __shared__ float s_out[];  // OUT_MODE 1 only
  __shared__ AxisBins s_by, s_bx;
  __shared__ RoiGeom s_g;

  const int r = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c0 = blockIdx.y * CH;
  const int c = c0 + lane * VEC;
  const int bins = PH * PW;

  if (threadIdx.x == 0) s_g = roi_geom(rois + 5 * (size_t)r, scale, sampling_ratio, aligned, PH, PW);
  for (int i = threadIdx.x; i < kMaxTaps; i += blockDim.x) { s_by.w[i] = 0.f; s_bx.w[i] = 0.f; }
  __syncthreads();
  const RoiGeom g = s_g;
  // dense windows need sample spacing <= 1 px (always true for the adaptive grid; not for a small fixed sampling_ratio)
  const bool tabled = PH <= 16 && PW <= 16 && PH * (g.gh + 1) <= kMaxTaps && PW * (g.gw + 1) <= kMaxTaps &&
                      g.bin_h <= (float)g.gh && g.bin_w <= (float)g.gw;
  if (tabled) {
    if ((int)threadIdx.x < PH) build_axis_bins(s_by, threadIdx.x, PH, g.gh, g.start_h, g.bin_h, H, W * C);
    else if (threadIdx.x >= 32 && (int)threadIdx.x - 32 < PW)
      build_axis_bins(s_bx, threadIdx.x - 32, PW, g.gw, g.start_w, g.bin_w, W, C);
  }
  __syncthreads();

  const T* fbase = feat + (size_t)g.batch * H * W * C + c;
  const bool active = c < C;
  const int row_stride = W * C;

  for (int b = warp; b < bins; b += kRoiWarps) {
    const int ph = b / PW, pw = b % PW;
    float acc[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) acc[k] = 0.f;
    if (active) {
      if (tabled) {
        const int ny = s_by.count[ph], nx = s_bx.count[pw];
        const float* wy = s_by.w + ph * (g.gh + 1);
        const float* wx = s_bx.w + pw * (g.gw + 1);
        const T* row = fbase + s_by.start[ph] + s_bx.start[pw];
        for (int ky = 0; ky < ny; ++ky, row += row_stride) {
          const float a = wy[ky];
          const T* p = row;
#pragma unroll 4
          for (int kx = 0; kx < nx; ++kx, p += C) {
            Vec<T, VEC> v;
            v.load(p);
            const float w = a * wx[kx];
#pragma unroll
            for (int k = 0; k < VEC; ++k) acc[k] += w * v.v[k];
          }
        }
      } else {
        for (int iy = 0; iy < g.gh; ++iy) {
          const AxisTap ty = make_tap(sample_coord(g.start_h, ph, g.bin_h, iy, g.gh), H, W * C);
          const T* row_lo = fbase + ty.lo;
          const T* row_hi = fbase + ty.hi;
          for (int ix = 0; ix < g.gw; ++ix) {
            const AxisTap tx = make_tap(sample_coord(g.start_w, pw, g.bin_w, ix, g.gw), W, C);
            Vec<T, VEC> v1, v2, v3, v4;
            v1.load(row_lo + tx.lo); v2.load(row_lo + tx.hi);
            v3.load(row_hi + tx.lo); v4.load(row_hi + tx.hi);
            const float w1 = ty.wlo * tx.wlo, w2 = ty.wlo * tx.whi, w3 = ty.whi * tx.wlo, w4 = ty.whi * tx.whi;
#pragma unroll
            for (int k = 0; k < VEC; ++k)
              acc[k] += w1 * v1.v[k] + w2 * v2.v[k] + w3 * v3.v[k] + w4 * v4.v[k];
          }
        }
      }
      // sample counts are small integers: 1/count is exact for powers of two and within 1 ulp otherwise
      const float inv = 1.0f / g.count;
#pragma unroll
      for (int k = 0; k < VEC; ++k) acc[k] *= inv;
    }
    if (OUT_MODE == 0) {
      if (active) Vec<T, VEC>::store_stream(out + ((size_t)r * bins + b) * C + c, acc);
    } else if (OUT_MODE == 1) {
#pragma unroll
      for (int k = 0; k < VEC; ++k) s_out[b * (CH + 1) + lane * VEC + k] = acc[k];
    } else {
      if (active) {
#pragma unroll
        for (int k = 0; k < VEC; ++k) out[((size_t)r * C + c + k) * bins + b] = from_float<T>(acc[k]);
      }
    }
  }
  if (OUT_MODE == 1) {
    __syncthreads();
    const int nch = min(CH, C - c0);
    T* obase = out + ((size_t)r * C + c0) * bins;  // contiguous [nch][bins] slab
    for (int i = threadIdx.x; i < nch * bins; i += blockDim.x) {
      const int cl = i / bins, b = i - cl * bins;
      obase[i] = from_float<T>(s_out[b * (CH + 1) + cl]);
    }
  }
}

template <typename T, int VEC>
static int launch_roi_fwd(const T* feat_nhwc, const float* rois, T* out, int C, int H, int W, int R, int PH, int PW,
                          float scale, int sr, int aligned, int out_layout, cudaStream_t st) {
  constexpr int CH = 32 * VEC;
  dim3 grid(R, ceil_div(C, CH)), block(kRoiWarps * 32);
  if (out_layout == B200_NHWC) {
    roi_align_fwd_nhwc_kernel<T, VEC, 0><<<grid, block, 0, st>>>(feat_nhwc, rois, out, C, H, W, PH, PW, scale, sr, aligned);
  } else {
    const size_t smem = (size_t)PH * PW * (CH + 1) * sizeof(float);
    if (smem <= 160 * 1024) {
      auto k = roi_align_fwd_nhwc_kernel<T, VEC, 1>;
      if (smem > 40 * 1024) B200_CUDA_CALL(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k<<<grid, block, smem, st>>>(feat_nhwc, rois, out, C, H, W, PH, PW, scale, sr, aligned);
    } else {
      roi_align_fwd_nhwc_kernel<T, VEC, 2><<<grid, block, 0, st>>>(feat_nhwc, rois, out, C, H, W, PH, PW, scale, sr, aligned);
    }
  }
  B200_CUDA_LAUNCH_CHECK("roi_align_fwd");
  return B200_OK;
}

}  // namespace b200

using namespace b200;

extern "C" size_t b200_roi_align_fwd_workspace_bytes(int N, int C, int H, int W, int dtype, int in_layout) {
  if (in_layout == B200_NHWC) return 0;
  return align_up((size_t)N * C * H * W * (dtype == B200_BF16 ? 2 : 4), 256);
}

extern "C" int b200_roi_align_fwd(const void* feat, const float* rois, void* out, int N, int C, int H, int W, int R,
                                  int pooled_h, int pooled_w, float spatial_scale, int sampling_ratio, int aligned,
                                  int dtype, int in_layout, int out_layout, void* workspace, size_t workspace_bytes,
                                  b200_stream_t stream) {
  B200_CHECK_ARG(R == 0 || (feat && out && rois), "roi_align_fwd: null tensor");
  B200_CHECK_ARG(N > 0 && C > 0 && H > 0 && W > 0 && R >= 0 && pooled_h > 0 && pooled_w > 0, "roi_align_fwd: bad shape");
  B200_CHECK_ARG((dtype | 1) == 1 && (in_layout | 1) == 1 && (out_layout | 1) == 1, "roi_align_fwd: bad dtype/layout");
  const int vec = dtype == B200_BF16 ? 8 : 4;
  if (C % vec != 0) {
    set_error("roi_align_fwd: C=%d must be a multiple of %d for this dtype", C, vec);
    return B200_ERR_UNSUPPORTED;
  }
  if ((size_t)H * W * C >= (1u << 30)) {
    set_error("roi_align_fwd: per-image map too large for 32-bit tap offsets");
    return B200_ERR_UNSUPPORTED;
  }
  if (R == 0) return B200_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const void* f = feat;
  if (in_layout == B200_NCHW) {
    const size_t need = b200_roi_align_fwd_workspace_bytes(N, C, H, W, dtype, in_layout);
    if (!workspace || workspace_bytes < need) {
      set_error("roi_align_fwd: workspace too small (%zu < %zu)", workspace_bytes, need);
      return B200_ERR_WORKSPACE;
    }
    int rc = dispatch_affine(feat, nullptr, nullptr, 1.0f, workspace, N, C, H, W, dtype, B200_NCHW, dtype, B200_NHWC, st);
    if (rc != B200_OK) return rc;
    f = workspace;
  }
  if (dtype == B200_F32)
    return launch_roi_fwd<float, 4>((const float*)f, rois, (float*)out, C, H, W, R, pooled_h, pooled_w, spatial_scale,
                                    sampling_ratio, aligned, out_layout, st);
  return launch_roi_fwd<__nv_bfloat16, 8>((const __nv_bfloat16*)f, rois, (__nv_bfloat16*)out, C, H, W, R, pooled_h,
                                          pooled_w, spatial_scale, sampling_ratio, aligned, out_layout, st);
}
