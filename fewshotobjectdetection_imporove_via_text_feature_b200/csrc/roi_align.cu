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
#include <cuda.h>
#include <string.h>

#include "common.cuh"
#include "roi_geom.cuh"

namespace b200 {

int dispatch_affine(const void* x, const float* w, const float* b, float mult, void* y, int N, int C, int H, int W,
                    int in_dtype, int in_layout, int out_dtype, int out_layout, cudaStream_t st);

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
                          int H, int W, int PH, int PW, int bin_step, float scale, int sampling_ratio, int aligned) {
  constexpr int CH = 32 * VEC;
  extern __shared__ float s_out[];  // OUT_MODE 1 only
  __shared__ AxisBins s_by, s_bx;
  __shared__ RoiGeom s_g;

  const int r = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c0 = blockIdx.y * CH;
  const int c = c0 + lane * VEC;
  // bin_step > 1: only bins (0, step, 2*step, ...) of each axis are computed and stored densely
  const int PWO = (PW + bin_step - 1) / bin_step;
  const int bins = ((PH + bin_step - 1) / bin_step) * PWO;

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
    const int ph = (b / PWO) * bin_step, pw = (b % PWO) * bin_step;
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

// ----------------------------------------------------------------------------------------------------------
// bf16 tensor-core variant.  ncu showed the CUDA-core kernel issue-bound (1 unpack + 1 FMA per channel-pixel), so
// for bf16 maps the per-bin-row contraction is put on the tensor cores:
//     out[ph][pw][c] = sum_px  W_ph[px][pw] * F[px][c],    W_ph[(y,x)][pw] = a_ph[y] * b_pw[x] / count
// as warp-level mma.sync m16n8k16 with  M = 16 channels, K = 16 pixels (one row segment), N = 8 bins (7 pw + pad).
// A fragments come straight from global memory: a lane loads 16 B (8 channels) of each of its 4 pixels and two
// PRMTs per channel pair transpose them into the (channel-row, pixel-pair) register layout — no smem staging; the
// row permutation (fragment rows g / g+8 <-> channels 2j, 2j+1 of the lane's 8) makes the D fragment hold 8
// consecutive channels per lane, so the epilogue is a plain 16-byte store.  B fragments (weights) are computed
// once per (row, 16-pixel tile) and reused by all 16 channel tiles of the warp.  CTA = (roi, 256 channels), one
// warp per output row ph.  Weights are rounded to bf16 (2^-9 relative), inside the bf16 output's own rounding.
// ----------------------------------------------------------------------------------------------------------

__device__ __forceinline__ void mma_bf16_16816(float* d, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                               uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack2_bf16(float lo, float hi) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&h);
}

// ----------------------------------------------------------------------------------------------------------
// TMA-fed version (the one dispatched for bf16 + channels-last output).  The register-fed variant above it in the
// history was correct but latency / L1TEX bound (ncu: 23 % issue-active, long-scoreboard stalls, L1TEX 75 %).
// Here every warp owns an output row ph and a private ring of shared-memory stages; one elected lane issues
// cp.async.bulk.tensor (4-D map over the NHWC feature map: {C, W, H, N}, box = 64 channels x 16 pixels of one
// row, 128B swizzle, OOB columns zero-filled) for the slabs ahead, completion on per-stage mbarriers, and the
// warp builds its A fragments with ldmatrix.x4.trans straight from the swizzled tiles (conflict-free).  The fp32
// D fragments are packed to bf16 and transposed back to channel-contiguous rows with stmatrix.trans, then leave
// as 16-byte coalesced stores.  No L1 involvement, loads in flight are bounded by smem, not by registers.
// ----------------------------------------------------------------------------------------------------------
constexpr int kTmaStages = 3;        // slabs in flight per warp
constexpr int kTmaBoxes = 2;         // 64-channel boxes per slab -> 128 channels per pass
constexpr int kBoxBytes = 16 * 128;  // 16 pixels x 64 bf16
constexpr int kWarpRingBytes = kTmaStages * kTmaBoxes * kBoxBytes;
constexpr int kWarpOutBytes = 8 * kTmaBoxes * 128;   // [8 pw][128 ch] bf16 staging for the epilogue

__device__ __forceinline__ uint32_t smem_addr_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tma_load_4d(uint32_t smem_dst, const CUtensorMap* map, uint32_t bar, int c, int x, int y,
                                            int n) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_dst), "l"(map), "r"(bar), "r"(c), "r"(x), "r"(y), "r"(n)
      : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_u32(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_u32(uint32_t bar, uint32_t parity) {
  const long long t0 = clock64();
  while (!mbar_try_wait_u32(bar, parity)) {
    if (clock64() - t0 > 4000000000ll) __trap();     // a protocol bug must fault, never hang the GPU
  }
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr) : "memory");
}
__device__ __forceinline__ void stmatrix_x4_trans(uint32_t addr, uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3) {
  asm volatile("stmatrix.sync.aligned.m8n8.x4.trans.shared.b16 [%0], {%1,%2,%3,%4};"
               ::"r"(addr), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
}

__global__ void __launch_bounds__(256, 2)
roi_align_fwd_tma_bf16_kernel(const __grid_constant__ CUtensorMap fmap, const __nv_bfloat16* __restrict__ feat,
                              const float* __restrict__ rois, __nv_bfloat16* __restrict__ out, int C, int H, int W,
                              int PH, int PW, float scale, int sampling_ratio, int aligned) {
  extern __shared__ unsigned char s_dyn_raw[];
  __shared__ AxisBins s_by, s_bx;
  __shared__ RoiGeom s_g;
  __shared__ int s_x0, s_nx;
  __shared__ __align__(8) unsigned long long s_bar[8][kTmaStages];

  const int r = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bins = PH * PW;
  unsigned char* s_dyn = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(s_dyn_raw) + 1023) & ~(uintptr_t)1023);

  if (threadIdx.x == 0) s_g = roi_geom(rois + 5 * (size_t)r, scale, sampling_ratio, aligned, PH, PW);
  for (int i = threadIdx.x; i < kMaxTaps; i += blockDim.x) { s_by.w[i] = 0.f; s_bx.w[i] = 0.f; }
  if (lane == 0 && warp < PH) {
    for (int st = 0; st < kTmaStages; ++st)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr_u32(&s_bar[warp][st])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const RoiGeom g = s_g;
  const bool tabled = PH * (g.gh + 1) <= kMaxTaps && PW * (g.gw + 1) <= kMaxTaps && g.bin_h <= (float)g.gh &&
                      g.bin_w <= (float)g.gw;
  if (tabled) {
    if ((int)threadIdx.x < PH) build_axis_bins(s_by, threadIdx.x, PH, g.gh, g.start_h, g.bin_h, H, 1);       // rows
    else if (threadIdx.x >= 32 && (int)threadIdx.x - 32 < PW)
      build_axis_bins(s_bx, threadIdx.x - 32, PW, g.gw, g.start_w, g.bin_w, W, 1);                           // columns
    if (threadIdx.x >= 32 + PW && threadIdx.x < 40) { s_bx.start[threadIdx.x - 32] = 0; s_bx.count[threadIdx.x - 32] = 0; }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int x0 = 1 << 30, xe = -1;
    for (int p = 0; p < PW; ++p)
      if (s_bx.count[p] > 0) { x0 = min(x0, s_bx.start[p]); xe = max(xe, s_bx.start[p] + s_bx.count[p]); }
    s_x0 = xe < 0 ? 0 : x0;
    s_nx = xe < 0 ? 0 : xe - x0;
  }
  __syncthreads();

  if (!tabled) {
    // rare shapes (fixed sparse sampling grids, very large windows): per-sample CUDA-core path
    const __nv_bfloat16* fimg = feat + (size_t)g.batch * H * W * C;
    const int nwarps = blockDim.x >> 5;
    for (int cc = 0; cc < C; cc += 256) {
      const int c = cc + lane * 8;
      for (int b = warp; b < bins; b += nwarps) {
        const int ph = b / PW, pw = b % PW;
        float acc[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = 0.f;
        if (c < C) {
          for (int iy = 0; iy < g.gh; ++iy) {
            const AxisTap ty = make_tap(sample_coord(g.start_h, ph, g.bin_h, iy, g.gh), H, W * C);
            for (int ix = 0; ix < g.gw; ++ix) {
              const AxisTap tx = make_tap(sample_coord(g.start_w, pw, g.bin_w, ix, g.gw), W, C);
              Vec<__nv_bfloat16, 8> v1, v2, v3, v4;
              v1.load(fimg + c + ty.lo + tx.lo); v2.load(fimg + c + ty.lo + tx.hi);
              v3.load(fimg + c + ty.hi + tx.lo); v4.load(fimg + c + ty.hi + tx.hi);
              const float w1 = ty.wlo * tx.wlo, w2 = ty.wlo * tx.whi, w3 = ty.whi * tx.wlo, w4 = ty.whi * tx.whi;
#pragma unroll
              for (int k = 0; k < 8; ++k) acc[k] += w1 * v1.v[k] + w2 * v2.v[k] + w3 * v3.v[k] + w4 * v4.v[k];
            }
          }
          const float inv = 1.0f / g.count;
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[k] *= inv;
          Vec<__nv_bfloat16, 8>::store_stream(out + ((size_t)r * bins + b) * C + c, acc);
        }
      }
    }
    return;
  }

  // ---- tensor-core path: warp <-> output row ph ---------------------------------------------------------------
  const int ph = warp;
  if (ph >= PH) return;
  const int gq = lane >> 2, t = lane & 3;
  const int nx = s_nx, x0 = s_x0, ny = s_by.count[ph], y0 = s_by.start[ph];
  const int ntx = (nx + 15) >> 4;
  const int npass = (C + 64 * kTmaBoxes - 1) / (64 * kTmaBoxes);
  const int slabs_per_pass = ntx * ny, total = npass * slabs_per_pass;
  const float inv_count = 1.0f / g.count;
  const float* wy = s_by.w + ph * (g.gh + 1);
  const int xrel = s_bx.start[gq] - x0, xcnt = s_bx.count[gq];
  const float* xwt = s_bx.w + gq * (g.gw + 1);
  auto xweight = [&](int k) -> float { const int i = k - xrel; return (i >= 0 && i < xcnt) ? xwt[i] : 0.f; };

  if (total == 0) {
    // every sample of this output row fell outside the map: the row is exactly zero
    for (int i = lane; i < PW * (C / 8); i += 32) {
      const int pw = i / (C / 8), c = (i - pw * (C / 8)) * 8;
      *reinterpret_cast<uint4*>(out + ((size_t)r * bins + ph * PW + pw) * C + c) = make_uint4(0, 0, 0, 0);
    }
    return;
  }
  unsigned char* ring = s_dyn + (size_t)warp * (kWarpRingBytes + kWarpOutBytes);
  const uint32_t ring_u32 = smem_addr_u32(ring);
  const uint32_t ostage_u32 = ring_u32 + kWarpRingBytes;
  const unsigned char* ostage = ring + kWarpRingBytes;
  const uint32_t bar0 = smem_addr_u32(&s_bar[warp][0]);

  // issue order == consume order: slab s -> (pass, tx, ky) with ky fastest
  int ip = 0, itx = 0, iky = 0, issued = 0;
  auto issue = [&]() {                       // lane 0 only
    const int st = issued % kTmaStages;
    const uint32_t bar = bar0 + 8 * st;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(kTmaBoxes * kBoxBytes) : "memory");
#pragma unroll
    for (int nb = 0; nb < kTmaBoxes; ++nb)
      tma_load_4d(ring_u32 + (st * kTmaBoxes + nb) * kBoxBytes, &fmap, bar, ip * 64 * kTmaBoxes + nb * 64, x0 + itx * 16,
                  y0 + iky, g.batch);
    ++issued;
    if (++iky == ny) { iky = 0; if (++itx == ntx) { itx = 0; ++ip; } }
  };
  if (lane == 0)
    for (int i = 0; i < kTmaStages && i < total; ++i) issue();

  // ldmatrix source address inside a box: lanes 0-7 / 8-15 / 16-23 / 24-31 address matrices 0..3 =
  // (px 0-7, chunk 2j) (px 0-7, chunk 2j+1) (px 8-15, chunk 2j) (px 8-15, chunk 2j+1); 128B swizzle: chunk ^= px & 7
  const int mi = lane >> 3, rr = lane & 7;
  const int lpx = (mi >> 1) * 8 + rr;
  const uint32_t lrow = lpx * 128;

  float acc[kTmaBoxes][4][4];
  int cky = 0, ctx = 0, cp = 0;
  float xw0 = 0.f, xw1 = 0.f, xw2 = 0.f, xw3 = 0.f;
  for (int s = 0; s < total; ++s) {
    if (cky == 0) {
      if (ctx == 0) {
#pragma unroll
        for (int i = 0; i < kTmaBoxes; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[i][j][k] = 0.f;
      }
      const int k = ctx * 16;
      xw0 = xweight(k + 2 * t); xw1 = xweight(k + 2 * t + 1); xw2 = xweight(k + 2 * t + 8); xw3 = xweight(k + 2 * t + 9);
    }
    const int st = s % kTmaStages;
    mbar_wait_u32(bar0 + 8 * st, (s / kTmaStages) & 1);
    const float a = wy[cky] * inv_count;
    const uint32_t b0 = pack2_bf16(a * xw0, a * xw1), b1 = pack2_bf16(a * xw2, a * xw3);
#pragma unroll
    for (int nb = 0; nb < kTmaBoxes; ++nb) {
      const uint32_t box = ring_u32 + (st * kTmaBoxes + nb) * kBoxBytes + lrow;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int chunk = 2 * j + (mi & 1);
        uint32_t a0, a1, a2, a3;
        ldmatrix_x4_trans(box + ((chunk ^ (lpx & 7)) << 4), a0, a1, a2, a3);
        mma_bf16_16816(acc[nb][j], a0, a1, a2, a3, b0, b1);
      }
    }
    __syncwarp();
    if (lane == 0 && issued < total) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic reads of this stage before the async refill
      issue();
    }
    if (++cky == ny) {
      cky = 0;
      if (++ctx == ntx) {
        ctx = 0;
        // ---- epilogue of pass cp: D (16 ch x 8 pw, fp32) -> bf16 -> stmatrix.trans -> [pw][128 ch] -> 16 B stores
#pragma unroll
        for (int nb = 0; nb < kTmaBoxes; ++nb) {
#pragma unroll
          for (int jj = 0; jj < 4; jj += 2) {
            // matrices: (tile jj, ch 0-7) (tile jj, ch 8-15) (tile jj+1, ch 0-7) (tile jj+1, ch 8-15); rows = channels,
            // cols = pw; .trans writes memory row (= lane's address) pw with 8 consecutive channels
            const uint32_t m0 = pack2_bf16(acc[nb][jj][0], acc[nb][jj][1]), m1 = pack2_bf16(acc[nb][jj][2], acc[nb][jj][3]);
            const uint32_t m2 = pack2_bf16(acc[nb][jj + 1][0], acc[nb][jj + 1][1]), m3 = pack2_bf16(acc[nb][jj + 1][2], acc[nb][jj + 1][3]);
            // lane (mi, rr): row rr (= pw) of matrix mi -> channel offset nb*64 + jj*16 + mi*8
            stmatrix_x4_trans(ostage_u32 + rr * (kTmaBoxes * 128) + (nb * 64 + jj * 16 + mi * 8) * 2, m0, m1, m2, m3);
          }
        }
        __syncwarp();
        const int c0 = cp * 64 * kTmaBoxes;
        constexpr int kLanesPerRow = kTmaBoxes * 8;            // 16-byte pieces per pw row
        for (int i = lane; i < PW * kLanesPerRow; i += 32) {
          const int pw = i / kLanesPerRow, piece = i - pw * kLanesPerRow;
          const int c = c0 + piece * 8;
          if (c < C) {
            const uint4 v = *reinterpret_cast<const uint4*>(ostage + pw * (kTmaBoxes * 128) + piece * 16);
            asm volatile("st.global.cs.v4.b32 [%0], {%1,%2,%3,%4};" ::"l"(out + ((size_t)r * bins + ph * PW + pw) * C + c),
                         "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
          }
        }
        __syncwarp();
        ++cp;
      }
    }
  }
}

template <typename T, int VEC>
static int launch_roi_fwd(const T* feat_nhwc, const float* rois, T* out, int C, int H, int W, int R, int PH, int PW,
                          int bin_step, float scale, int sr, int aligned, int out_layout, cudaStream_t st) {
  constexpr int CH = 32 * VEC;
  dim3 grid(R, ceil_div(C, CH)), block(kRoiWarps * 32);
  if (out_layout == B200_NHWC) {
    roi_align_fwd_nhwc_kernel<T, VEC, 0><<<grid, block, 0, st>>>(feat_nhwc, rois, out, C, H, W, PH, PW, bin_step, scale, sr, aligned);
  } else {
    const size_t smem = (size_t)ceil_div(PH, bin_step) * ceil_div(PW, bin_step) * (CH + 1) * sizeof(float);
    if (smem <= 160 * 1024) {
      auto k = roi_align_fwd_nhwc_kernel<T, VEC, 1>;
      if (smem > 40 * 1024) B200_CUDA_CALL(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k<<<grid, block, smem, st>>>(feat_nhwc, rois, out, C, H, W, PH, PW, bin_step, scale, sr, aligned);
    } else {
      roi_align_fwd_nhwc_kernel<T, VEC, 2><<<grid, block, 0, st>>>(feat_nhwc, rois, out, C, H, W, PH, PW, bin_step, scale, sr, aligned);
    }
  }
  B200_CUDA_LAUNCH_CHECK("roi_align_fwd");
  return B200_OK;
}

// 0: CUDA-core per-bin-window kernel (0.36 ms at R=4096, C=1024)
// 1: TMA + ldmatrix + mma.sync tensor-core kernel, per-ROI windows (0.50 ms: 2.5x fewer instructions, but it moves
//    3.4 GB L2->SM because windows are padded to 16-pixel boxes and rows shared by adjacent bins are fetched per bin row)
// 2: slice-resident tensor-core kernel (roi_align_slice.cu; default) — taken when the ROIs come grouped by image
//    (roi_batch_offsets given), bf16 channels-last in and out, 7x7 bins; otherwise the call falls back to 0
int g_roi_bf16_impl = 2;

extern int g_roi_bwd_impl;
size_t roi_slice_workspace_bytes(int R);
bool roi_slice_eligible(int C, int H, int W, int PH, int PW, int bin_step, const void* feat);
int launch_roi_fwd_slice(const __nv_bfloat16* feat, const float* rois, const int32_t* roi_offsets, __nv_bfloat16* out,
                         int N, int C, int H, int W, int R, int PH, int PW, int bin_step, float scale, int sr,
                         int aligned, void* workspace, cudaStream_t st);

typedef CUresult (*PFN_encodeTiledRoi)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                       const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                       CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int make_feature_map(CUtensorMap* m, const void* ptr, int N, int C, int H, int W) {
  static PFN_encodeTiledRoi enc = nullptr;
  if (!enc) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
      set_error("roi_align_fwd: cuTensorMapEncodeTiled unavailable");
      return B200_ERR_CUDA;
    }
    enc = (PFN_encodeTiledRoi)p;
  }
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {64, 16, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("roi_align_fwd: cuTensorMapEncodeTiled failed (%d)", (int)r);
    return B200_ERR_CUDA;
  }
  return B200_OK;
}

}  // namespace b200

namespace b200 { extern int g_gemm_ctas; extern int g_gemm_generic_epilogue; extern int g_roi_bwd_tile_variant; }
using namespace b200;

extern "C" int b200_set_option(const char* key, int value) {
  B200_CHECK_ARG(key != nullptr, "set_option: null key");
  if (strcmp(key, "roi_align_bf16_impl") == 0) {
    B200_CHECK_ARG(value >= 0 && value <= 2, "set_option: roi_align_bf16_impl must be 0 (cuda-core), 1 (tma+mma) or 2 (slice-resident)");
    g_roi_bf16_impl = value;
    return B200_OK;
  }
  if (strcmp(key, "roi_align_bwd_impl") == 0) {
    B200_CHECK_ARG(value >= 0 && value <= 2,
                   "set_option: roi_align_bwd_impl must be 0 (fp32 tables), 1 (per-pixel CSR gather) or 2 (pixel-tile tensor-core gather)");
    g_roi_bwd_impl = value;
    return B200_OK;
  }
  if (strcmp(key, "roi_bwd_tile_variant") == 0) {
    B200_CHECK_ARG(value >= 0 && value <= 3, "set_option: roi_bwd_tile_variant must be in [0, 3]");
    g_roi_bwd_tile_variant = value;
    return B200_OK;
  }
  if (strcmp(key, "gemm_generic_epilogue") == 0) {
    g_gemm_generic_epilogue = value != 0;
    return B200_OK;
  }
  if (strcmp(key, "gemm_ctas") == 0) {
    B200_CHECK_ARG(value >= 0 && value <= 4096, "set_option: gemm_ctas must be in [0, 4096] (0 = one persistent CTA per SM)");
    g_gemm_ctas = value;
    return B200_OK;
  }
  set_error("set_option: unknown key '%s'", key);
  return B200_ERR_INVALID;
}

extern "C" size_t b200_roi_align_fwd_workspace_bytes(int N, int C, int H, int W, int R, int dtype, int in_layout) {
  size_t b = 0;
  if (in_layout == B200_NCHW) b += align_up((size_t)N * C * H * W * (dtype == B200_BF16 ? 2 : 4), 256);
  if (dtype == B200_BF16) b += roi_slice_workspace_bytes(R);   // per-ROI geometry records of the slice-resident kernel
  return b;
}

extern "C" int b200_roi_align_fwd(const void* feat, const float* rois, const int32_t* roi_batch_offsets, void* out, int N,
                                  int C, int H, int W, int R, int pooled_h, int pooled_w, int bin_step,
                                  float spatial_scale, int sampling_ratio, int aligned, int dtype, int in_layout,
                                  int out_layout, void* workspace, size_t workspace_bytes, b200_stream_t stream) {
  B200_CHECK_ARG(R == 0 || (feat && out && rois), "roi_align_fwd: null tensor");
  B200_CHECK_ARG(N > 0 && C > 0 && H > 0 && W > 0 && R >= 0 && pooled_h > 0 && pooled_w > 0, "roi_align_fwd: bad shape");
  B200_CHECK_ARG((dtype | 1) == 1 && (in_layout | 1) == 1 && (out_layout | 1) == 1, "roi_align_fwd: bad dtype/layout");
  B200_CHECK_ARG(bin_step >= 1 && bin_step <= 8, "roi_align_fwd: bin_step must be in [1, 8]");
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
  const size_t need = b200_roi_align_fwd_workspace_bytes(N, C, H, W, R, dtype, in_layout);
  if (need && (!workspace || workspace_bytes < need)) {
    set_error("roi_align_fwd: workspace too small (%zu < %zu)", workspace_bytes, need);
    return B200_ERR_WORKSPACE;
  }
  unsigned char* ws = (unsigned char*)workspace;
  const void* f = feat;
  if (in_layout == B200_NCHW) {
    int rc = dispatch_affine(feat, nullptr, nullptr, 1.0f, ws, N, C, H, W, dtype, B200_NCHW, dtype, B200_NHWC, st);
    if (rc != B200_OK) return rc;
    f = ws;
    ws += align_up((size_t)N * C * H * W * (dtype == B200_BF16 ? 2 : 4), 256);
  }
  if (dtype == B200_F32)
    return launch_roi_fwd<float, 4>((const float*)f, rois, (float*)out, C, H, W, R, pooled_h, pooled_w, bin_step,
                                    spatial_scale, sampling_ratio, aligned, out_layout, st);
  if (g_roi_bf16_impl == 2 && roi_batch_offsets && out_layout == B200_NHWC &&
      roi_slice_eligible(C, H, W, pooled_h, pooled_w, bin_step, f))
    return launch_roi_fwd_slice((const __nv_bfloat16*)f, rois, roi_batch_offsets, (__nv_bfloat16*)out, N, C, H, W, R,
                                pooled_h, pooled_w, bin_step, spatial_scale, sampling_ratio, aligned, ws, st);
  if (g_roi_bf16_impl == 1 && bin_step == 1 && out_layout == B200_NHWC && pooled_h <= 8 && pooled_w <= 8 && ((uintptr_t)f & 15) == 0) {
    // tensor-core path: 4-D TMA map over the NHWC map {C, W, H, N}, box 64 ch x 16 px, 128B swizzle
    CUtensorMap fmap;
    int rc = make_feature_map(&fmap, f, N, C, H, W);
    if (rc != B200_OK) return rc;
    const int warps = max(2, pooled_h);
    const size_t smem = (size_t)pooled_h * (kWarpRingBytes + kWarpOutBytes) + 1024;
    // set on every call: the attribute is per device and the call is cheap (no process-wide flag to race on)
    B200_CUDA_CALL(cudaFuncSetAttribute(roi_align_fwd_tma_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * (kWarpRingBytes + kWarpOutBytes) + 1024));
    roi_align_fwd_tma_bf16_kernel<<<R, 32 * warps, smem, st>>>(fmap, (const __nv_bfloat16*)f, rois, (__nv_bfloat16*)out, C, H,
                                                               W, pooled_h, pooled_w, spatial_scale, sampling_ratio, aligned);
    B200_CUDA_LAUNCH_CHECK("roi_align_fwd_tma");
    return B200_OK;
  }
  return launch_roi_fwd<__nv_bfloat16, 8>((const __nv_bfloat16*)f, rois, (__nv_bfloat16*)out, C, H, W, R, pooled_h,
                                          pooled_w, bin_step, spatial_scale, sampling_ratio, aligned, out_layout, st);
}
