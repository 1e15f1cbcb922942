// P1b: ROIAlign backward, row-gather bf16 kernel — atomic-free, deterministic, the fine-tune path's default.
// Reference: autograd of the roi_align call at defrcn/modeling/roi_heads/roi_heads.py:340 (torchvision scatters with
// atomicAdd; fine-tuning reaches it with BACKWARD_SCALE = 0.001 through the GDL).
//
// Gather form of the separable identity (see roi_align_bwd.cu), driven by the per-ROI geometry records that the
// forward's prepare kernel builds (roi_align_slice.cu):
//   grad_feat[n,y,x,c] = sum over ROIs r of image n (index order), output rows ph whose window contains y,
//                        output columns pw whose window contains x:  a_r,ph[y] * b_r,pw[x] * g[r,ph,pw,c]
// CTA = (map row y, image n, 128 channels).  Phase 1 compacts, in ROI order, the (roi, ph, a) pairs that touch row y
// into shared memory (one warp, ballot/scan).  Phase 2 gives every (pixel, 8-channel group) of the row a thread that
// walks the list and accumulates in registers — each output element has exactly one writer and a fixed summation
// order, so the result is bitwise reproducible and nothing is zero-filled or atomically updated.
#include "common.cuh"
#include "roi_geom.cuh"
#include "roi_slice_rec.cuh"

namespace b200 {

constexpr int kGatherLanes = 16;      // threads per pixel: 16 x 8 channels = 128 channels per CTA
constexpr int kGatherMaxPx = 64;      // pixels per pass of the CTA
constexpr int kGatherMaxPass = 4;     // map width <= 256 (2 passes up to 64 columns: smaller CTAs, more of them per SM)
constexpr int kGatherCap = 2048;      // row-list entries held in shared memory at a time (longer lists go in rounds)
constexpr int kGatherMaxGroups = 16;  // 32-ROI groups scanned per round, one warp each

struct RowEntry {
  int roi;
  int pho;        // computed output row (index into the strided gradient), -1: table-less ROI (per-sample path)
  float a;        // vertical weight a_ph[y] / count (bf16-rounded, as the forward uses it)
  int xext;       // xlo | xhi << 16 : pixel columns touched by the computed bins
};

__global__ void __launch_bounds__(kGatherLanes* kGatherMaxPx)
roi_align_bwd_gather_kernel(const __nv_bfloat16* __restrict__ g, const unsigned char* __restrict__ recs,
                            const float* __restrict__ rois, const int32_t* __restrict__ roi_offsets,
                            __nv_bfloat16* __restrict__ grad_feat, int C, int H, int W, int PH, int PW, int bin_step,
                            float scale, int sampling_ratio, int aligned) {
  __shared__ RowEntry s_ent[kGatherCap];
  __shared__ int s_n, s_next, s_gtot[kGatherMaxGroups];
  const int y = blockIdx.x, n = blockIdx.y;
  const int r0 = roi_offsets[n], r1 = roi_offsets[n + 1];
  const int PHO = (PH + bin_step - 1) / bin_step, PWO = (PW + bin_step - 1) / bin_step;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const int sub = threadIdx.x & (kGatherLanes - 1);
  const int c = blockIdx.z * (kGatherLanes * 8) + sub * 8;
  const int px_per_pass = blockDim.x / kGatherLanes;
  const int x0 = threadIdx.x / kGatherLanes;
  float acc[kGatherMaxPass][8];
#pragma unroll
  for (int p = 0; p < kGatherMaxPass; ++p)
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[p][k] = 0.f;
  if (threadIdx.x == 0) s_next = r0;
  __syncthreads();

  for (;;) {
  const int rstart = s_next;
  if (rstart >= r1) break;
  __syncthreads();                                   // everyone has read s_next / finished the previous round's list
  // ---- phase 1: ordered list of the (roi, ph) windows that contain row y.  Each warp scans one group of 32 ROIs
  // (lane <-> ROI); group totals are exchanged through shared memory so that the list keeps the ROI index order.
  const int nwarps = blockDim.x >> 5;
  const int ngroups = min(min(nwarps, kGatherMaxGroups), min((r1 - rstart + 31) >> 5, kGatherCap / (32 * PHO)));
  int cnt = 0, first_pho = 0, flags = 0, xext = 0, incl = 0;
  const int r = rstart + warp * 32 + lane;
  const unsigned char* rec = recs + (size_t)min(r, r1 - 1) * kRecBytes;
  if (warp < ngroups) {
    if (r < r1) {
      flags = reinterpret_cast<const int*>(rec)[kOffFlags / 4];
      if (!flags) {
        cnt = 1;                                  // table-less ROI: one entry, resolved per sample in phase 2
      } else {
        const int ylo = reinterpret_cast<const int*>(rec)[kOffYExt / 4], yhi = reinterpret_cast<const int*>(rec)[kOffYExt / 4 + 1];
        if (y >= ylo && y <= yhi) {
          const uint2 ysb = *reinterpret_cast<const uint2*>(rec + kOffYStart);
          const uint2 ycb = *reinterpret_cast<const uint2*>(rec + kOffYCount);
          bool seen = false;
          for (int pho = 0; pho < PHO; ++pho) {
            const int ph = pho * bin_step;
            const int ys = ((ph < 4 ? ysb.x : ysb.y) >> (8 * (ph & 3))) & 255;
            const int yc = ((ph < 4 ? ycb.x : ycb.y) >> (8 * (ph & 3))) & 255;
            if (y >= ys && y < ys + yc) {         // windows containing y are consecutive in ph
              if (!seen) { first_pho = pho; seen = true; }
              ++cnt;
            }
          }
          if (cnt) {
            int xlo = 1 << 15, xhi = -1;
            for (int pw = 0; pw < PW; pw += bin_step) {
              const int xs = rec[kOffXStart + pw], xc = rec[kOffXCount + pw];
              if (xc) { xlo = min(xlo, xs); xhi = max(xhi, xs + xc - 1); }
            }
            if (xhi < 0) cnt = 0;
            xext = xlo | (xhi << 16);
          }
        }
      }
    }
    incl = cnt;                                   // inclusive warp scan of the per-ROI entry counts
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) s_gtot[warp] = incl;
  }
  __syncthreads();
  if (warp < ngroups) {
    int pos = incl - cnt;
    for (int i = 0; i < warp; ++i) pos += s_gtot[i];
    if (cnt) {
      if (!flags) {
        s_ent[pos] = RowEntry{r, -1, 0.f, 0};
      } else {
        const uint32_t* wy2 = reinterpret_cast<const uint32_t*>(rec + kOffWy);
        for (int i = 0; i < cnt; ++i) {
          const int pho = first_pho + i, ph = pho * bin_step;
          const int ys = rec[kOffYStart + ph];
          s_ent[pos + i] = RowEntry{r, pho, __uint_as_float(wy2[ph * kTaps + (y - ys)] << 16), xext};
        }
      }
    }
  }
  if (threadIdx.x == 0) {
    int total = 0;
    for (int i = 0; i < ngroups; ++i) total += s_gtot[i];
    s_n = total;
    s_next = rstart + ngroups * 32;
  }
  __syncthreads();
  const int nent = s_n;

  // ---- phase 2: one thread per (pixel, 8 channels), accumulators live in registers across the rounds -------------
#pragma unroll
  for (int p = 0; p < kGatherMaxPass; ++p) {
    const int x = x0 + p * px_per_pass;
    if (x < W && c < C) {
      float* ac = acc[p];
      for (int e = 0; e < nent; ++e) {
        const RowEntry en = s_ent[e];
        if (en.pho >= 0) {
          if (x < (en.xext & 0xffff) || x > (en.xext >> 16)) continue;
          const unsigned char* rec = recs + (size_t)en.roi * kRecBytes;
          const uint2 xsb = __ldg(reinterpret_cast<const uint2*>(rec + kOffXStart));
          const uint2 xcb = __ldg(reinterpret_cast<const uint2*>(rec + kOffXCount));
          const float* wx = reinterpret_cast<const float*>(rec + kOffWx);
          const __nv_bfloat16* grow = g + ((size_t)en.roi * PHO + en.pho) * PWO * C + c;
          for (int pwo = 0; pwo < PWO; ++pwo) {
            const int pw = pwo * bin_step;
            const int xs = ((pw < 4 ? xsb.x : xsb.y) >> (8 * (pw & 3))) & 255;
            const int xc = ((pw < 4 ? xcb.x : xcb.y) >> (8 * (pw & 3))) & 255;
            const int k = x - xs;
            if ((unsigned)k < (unsigned)xc) {
              const float w = en.a * __ldg(wx + pw * kTaps + k);
              const uint4 t = __ldg(reinterpret_cast<const uint4*>(grow + (size_t)pwo * C));
              ac[0] += w * __uint_as_float(t.x << 16); ac[1] += w * __uint_as_float(t.x & 0xffff0000u);
              ac[2] += w * __uint_as_float(t.y << 16); ac[3] += w * __uint_as_float(t.y & 0xffff0000u);
              ac[4] += w * __uint_as_float(t.z << 16); ac[5] += w * __uint_as_float(t.z & 0xffff0000u);
              ac[6] += w * __uint_as_float(t.w << 16); ac[7] += w * __uint_as_float(t.w & 0xffff0000u);
            }
          }
        } else {
          // rare shapes (sparse fixed sampling grids, windows wider than the tables): per-sample taps
          const RoiGeom q = roi_geom(rois + 5 * (size_t)en.roi, scale, sampling_ratio, aligned, PH, PW);
          const float inv = 1.0f / q.count;
          for (int pho = 0; pho < PHO; ++pho)
            for (int iy = 0; iy < q.gh; ++iy) {
              const AxisTap ty = make_tap(sample_coord(q.start_h, pho * bin_step, q.bin_h, iy, q.gh), H, 1);
              const float wyv = (ty.lo == y ? ty.wlo : 0.f) + (ty.hi == y ? ty.whi : 0.f);
              if (wyv == 0.f) continue;
              for (int pwo = 0; pwo < PWO; ++pwo) {
                float wxv = 0.f;
                for (int ix = 0; ix < q.gw; ++ix) {
                  const AxisTap tx = make_tap(sample_coord(q.start_w, pwo * bin_step, q.bin_w, ix, q.gw), W, 1);
                  wxv += (tx.lo == x ? tx.wlo : 0.f) + (tx.hi == x ? tx.whi : 0.f);
                }
                if (wxv == 0.f) continue;
                const float w = wyv * wxv * inv;
                const uint4 t = __ldg(reinterpret_cast<const uint4*>(g + (((size_t)en.roi * PHO + pho) * PWO + pwo) * C + c));
                ac[0] += w * __uint_as_float(t.x << 16); ac[1] += w * __uint_as_float(t.x & 0xffff0000u);
                ac[2] += w * __uint_as_float(t.y << 16); ac[3] += w * __uint_as_float(t.y & 0xffff0000u);
                ac[4] += w * __uint_as_float(t.z << 16); ac[5] += w * __uint_as_float(t.z & 0xffff0000u);
                ac[6] += w * __uint_as_float(t.w << 16); ac[7] += w * __uint_as_float(t.w & 0xffff0000u);
              }
            }
        }
      }
    }
  }
  }   // rounds
#pragma unroll
  for (int p = 0; p < kGatherMaxPass; ++p) {
    const int x = x0 + p * px_per_pass;
    if (x < W && c < C) {
      const float* ac = acc[p];
      uint4 o;
      __nv_bfloat162 h0 = __floats2bfloat162_rn(ac[0], ac[1]), h1 = __floats2bfloat162_rn(ac[2], ac[3]);
      __nv_bfloat162 h2 = __floats2bfloat162_rn(ac[4], ac[5]), h3 = __floats2bfloat162_rn(ac[6], ac[7]);
      o.x = *reinterpret_cast<uint32_t*>(&h0); o.y = *reinterpret_cast<uint32_t*>(&h1);
      o.z = *reinterpret_cast<uint32_t*>(&h2); o.w = *reinterpret_cast<uint32_t*>(&h3);
      *reinterpret_cast<uint4*>(grad_feat + (((size_t)n * H + y) * W + x) * C + c) = o;
    }
  }
}

bool roi_bwd_slice_eligible(int C, int H, int W, int PH, int PW, int bin_step) {
  return PH <= 7 && PW <= 7 && bin_step >= 1 && C % 8 == 0 && H <= 256 && W <= 256;
}

size_t roi_bwd_slice_workspace_bytes(int R) { return align_up((size_t)max(R, 1) * kRecBytes, 256); }

int launch_roi_bwd_slice(const __nv_bfloat16* g, const float* rois, const int32_t* roi_offsets, __nv_bfloat16* grad_feat,
                         int N, int C, int H, int W, int R, int PH, int PW, int bin_step, float scale, int sr, int aligned,
                         void* workspace, cudaStream_t st) {
  unsigned char* recs = (unsigned char*)workspace;
  int rc = launch_roi_slice_prepare(rois, recs, R, H, W, PH, PW, bin_step, scale, sr, aligned, st);
  if (rc != B200_OK) return rc;
  const int passes = W <= 64 ? 2 : 4;
  const int px = (max(2, ceil_div(W, passes)) + 1) & ~1;      // even: whole warps
  dim3 grid(H, N, ceil_div(C, kGatherLanes * 8));
  roi_align_bwd_gather_kernel<<<grid, px * kGatherLanes, 0, st>>>(g, recs, rois, roi_offsets, grad_feat, C, H, W, PH, PW,
                                                                 bin_step, scale, sr, aligned);
  B200_CUDA_LAUNCH_CHECK("roi_align_bwd_gather");
  return B200_OK;
}

}  // namespace b200
