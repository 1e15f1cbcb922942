// P1: ROIAlign forward, "slice-resident" bf16 kernel — the default for the head's 7x7 @ 1/16 pooling.
// Reference call site: defrcn/modeling/roi_heads/roi_heads.py:300-305,339-340 (detectron2 ROIPooler -> torchvision
// roi_align, aligned=True, adaptive sampling grid).
//
// Why: the per-ROI kernels (roi_align.cu) re-read every ROI window from L2 (1.8 GB L2->SM per launch for 0.44 GB of
// algorithmic bytes) and spend >2 issue slots per channel-pixel on CUDA cores.  Here a persistent CTA keeps a
// 32-channel slice of one image's whole feature map resident in shared memory (TMA, 64B swizzle, row pitch padded
// to a multiple of 8 pixels so the swizzle term does not depend on the row) and sweeps that image's ROIs:
//   * HBM/L2 -> SM traffic is the map once per (CTA, slice) + a 640-byte geometry record per (ROI, slice);
//   * the contraction runs on the tensor cores: out[ch16 x pw8] += F^T[ch16 x px8] * Wgt[px8 x pw8] as
//     mma.sync m16n8k8 (bf16, fp32 accumulate).  A fragments come from ldmatrix.x4.trans on the swizzled slice (8
//     pixels x 32 channels per instruction, conflict-free), B fragments are the separable bilinear weights
//     bf16(a_ph[y] / count) * bf16(b_pw[x]) (one packed multiply per fragment; 3 bf16 roundings, < 0.6 % relative).  One warp owns one ROI at a time; all PHO output rows
//     accumulate in registers, the epilogue transposes through stmatrix and leaves as 16-byte streaming stores.
//   * geometry (per-bin pixel windows and summed weights, exactly the tables of roi_align.cu) is computed once per
//     ROI by a tiny prepare kernel instead of once per (ROI, channel chunk).
// `bin_step` = 2 computes only the bins (0,2,4,6) x (0,2,4,6): res5's first block reads the pooled map through 1x1
// stride-2 convolutions (roi_heads.py:313-337 with RESNETS.STRIDE_IN_1X1), so the other 33 of 49 bins are dead.
#include <cuda.h>

#include "common.cuh"
#include "roi_geom.cuh"
#include "roi_slice_rec.cuh"

namespace b200 {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  const long long t0 = clock64();
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return;
    if (clock64() - t0 > 4000000000ll) __trap();   // a protocol bug must fault, never hang the GPU
  }
}

__device__ __forceinline__ void mma_bf16_1688(float* d, uint32_t a0, uint32_t a1, uint32_t b0) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(b0));
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&h);
}

// ---- prepare: one warp per ROI builds its record -----------------------------------------------------------------
struct AxisWin {
  int first, cnt;
  float w[kTaps];
};

__device__ __forceinline__ void build_win(AxisWin& o, int p, int g, float start, float bin, int size) {
  o.first = -1; o.cnt = 0;
#pragma unroll
  for (int k = 0; k < kTaps; ++k) o.w[k] = 0.f;
  for (int i = 0; i < g; ++i) {
    float coord = sample_coord(start, p, bin, i, g);
    if (coord < -1.0f || coord > (float)size) continue;
    if (coord <= 0.f) coord = 0.f;
    int lo = (int)coord, hi;
    if (lo >= size - 1) { hi = lo = size - 1; coord = (float)lo; } else hi = lo + 1;
    const float l = coord - (float)lo;
    if (o.first < 0) o.first = lo;
    const int klo = lo - o.first, khi = hi - o.first;   // < kTaps: sample spacing <= 1 px and g + 1 <= kTaps
#pragma unroll
    for (int k = 0; k < kTaps; ++k) {                   // static indexing keeps w[] in registers
      if (k == klo) o.w[k] += 1.f - l;
      if (k == khi) o.w[k] += l;
    }
    o.cnt = max(o.cnt, khi + 1);
  }
  o.first = max(o.first, 0);
}

// B-fragment weights of lane (g = lane >> 2, t = lane & 3) for pixel-list entries ia, ia + 1 and bin pw
__device__ __forceinline__ uint32_t xw_pair(const unsigned char* rb, int ia, int nxs, int xs_lo, int xs_n,
                                            const float* wxp) {
  float xwa = 0.f, xwb = 0.f;
  if (ia < nxs) {
    const int k = (int)rb[kOffXs + ia] - xs_lo;
    if ((unsigned)k < (unsigned)xs_n) xwa = wxp[k];
  }
  if (ia + 1 < nxs) {
    const int k = (int)rb[kOffXs + ia + 1] - xs_lo;
    if ((unsigned)k < (unsigned)xs_n) xwb = wxp[k];
  }
  return pack_bf16x2(xwa, xwb);
}

__global__ void __launch_bounds__(256)
roi_slice_prepare_kernel(const float* __restrict__ rois, unsigned char* __restrict__ recs, int R, int H, int W, int PH,
                         int PW, int bin_step, float scale, int sampling_ratio, int aligned) {
  __shared__ __align__(16) unsigned char s_rec[8][kRecBytes];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * 8 + warp;
  if (r >= R) return;
  unsigned char* rec = s_rec[warp];
  for (int i = lane; i < kRecBytes / 4; i += 32) reinterpret_cast<uint32_t*>(rec)[i] = 0u;
  __syncwarp();
  const RoiGeom g = roi_geom(rois + 5 * (size_t)r, scale, sampling_ratio, aligned, PH, PW);
  // the table path needs dense windows (sample spacing <= 1 px) of at most kTaps pixels per bin and axis
  bool ok = g.gh + 1 <= kTaps && g.gw + 1 <= kTaps && g.bin_h <= (float)g.gh && g.bin_w <= (float)g.gw;
  const float inv = 1.0f / g.count;
  const bool is_y = lane < PH, is_x = lane >= 8 && lane - 8 < PW;
  AxisWin win;
  win.first = 0; win.cnt = 0;
  if (ok && is_y) build_win(win, lane, g.gh, g.start_h, g.bin_h, H);
  if (ok && is_x) build_win(win, lane - 8, g.gw, g.start_w, g.bin_w, W);
  if (ok && is_y) {
    rec[kOffYStart + lane] = (unsigned char)win.first;
    rec[kOffYCount + lane] = (unsigned char)win.cnt;
    uint32_t* wy = reinterpret_cast<uint32_t*>(rec + kOffWy) + lane * kTaps;
#pragma unroll
    for (int k = 0; k < kTaps; ++k) { const float a = win.w[k] * inv; wy[k] = pack_bf16x2(a, a); }
  }
  if (ok && is_x) {
    rec[kOffXStart + lane - 8] = (unsigned char)win.first;
    rec[kOffXCount + lane - 8] = (unsigned char)win.cnt;
    float* wx = reinterpret_cast<float*>(rec + kOffWx) + (lane - 8) * kTaps;
#pragma unroll
    for (int k = 0; k < kTaps; ++k) wx[k] = win.w[k];
  }
  __syncwarp();
  if (lane == 0) {
    // sorted union of the pixel columns of the bins that are computed (window starts are non-decreasing in pw)
    int n = 0, last = -1;
    if (ok) {
      for (int pw = 0; pw < PW && ok; pw += bin_step) {
        const int s = rec[kOffXStart + pw], c = rec[kOffXCount + pw];
        for (int k = 0; k < c; ++k) {
          const int x = s + k;
          if (x <= last) continue;
          if (n == kMaxXs) { ok = false; break; }
          rec[kOffXs + n++] = (unsigned char)x;
          last = x;
        }
      }
    }
    int nymax = 0, ylo = 1 << 20, yhi = -1;
    for (int ph = 0; ph < PH; ph += bin_step) {
      const int yc = rec[kOffYCount + ph];
      nymax = max(nymax, yc);
      if (yc > 0) { ylo = min(ylo, (int)rec[kOffYStart + ph]); yhi = max(yhi, (int)rec[kOffYStart + ph] + yc - 1); }
    }
    reinterpret_cast<int*>(rec)[kOffYExt / 4] = ylo;
    reinterpret_cast<int*>(rec)[kOffYExt / 4 + 1] = yhi;
    int xlo = 1 << 20, xhi = -1;
    for (int pw = 0; pw < PW; pw += bin_step) {
      const int xc = rec[kOffXCount + pw];
      if (xc > 0) { xlo = min(xlo, (int)rec[kOffXStart + pw]); xhi = max(xhi, (int)rec[kOffXStart + pw] + xc - 1); }
    }
    reinterpret_cast<int*>(rec)[kOffXExt / 4] = xlo;
    reinterpret_cast<int*>(rec)[kOffXExt / 4 + 1] = xhi;
    reinterpret_cast<int*>(rec)[kOffBatch / 4] = g.batch;
    reinterpret_cast<int*>(rec)[kOffFlags / 4] = ok ? 1 : 0;
    reinterpret_cast<int*>(rec)[kOffNxs / 4] = ok ? n : 0;
    reinterpret_cast<int*>(rec)[kOffNyMax / 4] = ok ? nymax : 0;
  }
  __syncwarp();
  if (reinterpret_cast<const int*>(rec)[kOffFlags / 4]) {
    // B-fragment weight pairs of the first kTabTiles tiles, exactly as the main kernel's lane (g,t) would build them
    const int nxs = reinterpret_cast<const int*>(rec)[kOffNxs / 4];
    const int gq = lane >> 2, t = lane & 3;
    const int pw = gq * bin_step;
    const bool pw_ok = pw < PW;
    const int xs_lo = pw_ok ? rec[kOffXStart + pw] : 0, xs_n = pw_ok ? rec[kOffXCount + pw] : 0;
    const float* wxp = reinterpret_cast<const float*>(rec + kOffWx) + (pw_ok ? pw : 0) * kTaps;
    for (int tile = 0; tile < kTabTiles; ++tile)
      reinterpret_cast<uint32_t*>(rec + kOffXw2)[tile * 32 + lane] = xw_pair(rec, tile * 8 + 2 * t, nxs, xs_lo, xs_n, wxp);
  }
  __syncwarp();
  uint4* dst = reinterpret_cast<uint4*>(recs + (size_t)r * kRecBytes);
  for (int i = lane; i < kRecChunks; i += 32) dst[i] = reinterpret_cast<const uint4*>(rec)[i];
}

// ---- main kernel ---------------------------------------------------------------------------------------------------
struct SegInfo {
  int n, s, roi_begin, len;
};

__device__ __forceinline__ void rec_prefetch(uint32_t dst, const unsigned char* src, int lane) {
#pragma unroll
  for (int i = 0; i < kRecChunks / 32; ++i)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + (i * 32 + lane) * 16), "l"(src + (i * 32 + lane) * 16)
                 : "memory");
}

template <int PHO, int PWO, int STEP>
__global__ void __launch_bounds__(kSliceWarps * 32, 1)
roi_align_fwd_slice_kernel(const __grid_constant__ CUtensorMap fmap, const unsigned char* __restrict__ recs,
                           const float* __restrict__ rois, const int32_t* __restrict__ roi_offsets,
                           __nv_bfloat16* __restrict__ out, int N, int C, int H, int W, int Wp, int PH, int PW,
                           float scale, int sampling_ratio, int aligned) {
  extern __shared__ unsigned char s_raw[];
  __shared__ __align__(8) unsigned long long s_bar;
  __shared__ SegInfo s_seg;
  __shared__ int s_ctr;

  unsigned char* s_slice = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(s_raw) + 1023) & ~(uintptr_t)1023);
  const int rowbytes = Wp * kPixBytes;
  const int slice_bytes = H * rowbytes;
  unsigned char* s_recs = s_slice + slice_bytes;                          // [warps][kRecRing][kRecBytes]
  unsigned char* s_stage = s_recs + kSliceWarps * kRecRing * kRecBytes;    // [warps][kStageBytes]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = C / kSliceCh;
  const int R = roi_offsets[N];
  const long long T = (long long)R * S;
  const long long q0 = T * blockIdx.x / gridDim.x, q1 = T * (blockIdx.x + 1) / gridDim.x;
  const uint32_t bar = smem_u32(&s_bar);
  const uint32_t slice_u32 = smem_u32(s_slice);

  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  unsigned char* my_recs = s_recs + warp * kRecRing * kRecBytes;
  const uint32_t my_recs_u32 = smem_u32(my_recs);
  unsigned char* my_stage = s_stage + warp * kStageBytes;
  const uint32_t my_stage_u32 = smem_u32(my_stage);
  const int gq = lane >> 2, t = lane & 3, mi = lane >> 3, rr = lane & 7;

  uint32_t phase = 0;
  long long q = q0;
  while (q < q1) {
    __syncthreads();                    // every warp is done reading the previous slice and segment info
    if (threadIdx.x == 0) {
      int n = 0;
      while (n + 1 < N && (long long)roi_offsets[n + 1] * S <= q) ++n;
      const int r0 = roi_offsets[n], cnt = roi_offsets[n + 1] - r0;
      const long long rel = q - (long long)r0 * S;
      const int s = (int)(rel / cnt), i = (int)(rel - (long long)s * cnt);
      SegInfo sg;
      sg.n = n; sg.s = s; sg.roi_begin = r0 + i;
      sg.len = (int)min((long long)(cnt - i), q1 - q);
      s_seg = sg;
      s_ctr = kSliceWarps;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic reads of the slice before the async refill
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(slice_bytes) : "memory");
      asm volatile(
          "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
          ::"r"(slice_u32), "l"(&fmap), "r"(bar), "r"(s * kSliceCh), "r"(0), "r"(0), "r"(n)
          : "memory");
    }
    __syncthreads();
    const SegInfo sg = s_seg;
    q += sg.len;
    const int c0 = sg.s * kSliceCh;

    // warp w starts on ROI w of the segment, further ROIs are handed out from a shared counter (the ROIs' costs differ
    // by >10x, a static split leaves warps idle at the segment end); the next record streams in while this one is used
    int idx = warp, buf = 0;
    if (idx < sg.len) rec_prefetch(my_recs_u32, recs + (size_t)(sg.roi_begin + idx) * kRecBytes, lane);
    asm volatile("cp.async.commit_group;" ::: "memory");
    int nidx = 0;
    if (lane == 0) nidx = atomicAdd(&s_ctr, 1);
    nidx = __shfl_sync(0xffffffffu, nidx, 0);
    mbar_wait(bar, phase);
    phase ^= 1;
    for (; idx < sg.len; buf ^= 1) {
      if (nidx < sg.len)
        rec_prefetch(my_recs_u32 + (buf ^ 1) * kRecBytes, recs + (size_t)(sg.roi_begin + nidx) * kRecBytes, lane);
      asm volatile("cp.async.commit_group;" ::: "memory");
      int nnidx = 0;
      if (lane == 0 && nidx < sg.len) nnidx = atomicAdd(&s_ctr, 1);   // consumed at the end of this iteration
      asm volatile("cp.async.wait_group 1;" ::: "memory");
      __syncwarp();
      const unsigned char* rb = my_recs + buf * kRecBytes;
      const float* rf = reinterpret_cast<const float*>(rb);
      const int roi = sg.roi_begin + idx;
      const int flags = reinterpret_cast<const int*>(rb)[kOffFlags / 4];
      __nv_bfloat16* obase = out + (size_t)roi * PHO * PWO * C + c0;

      if (flags) {
        const int nxs = reinterpret_cast<const int*>(rb)[kOffNxs / 4];
        const int nymax = reinterpret_cast<const int*>(rb)[kOffNyMax / 4];
        const int ntiles = (nxs + 7) >> 3;
        float acc[PHO][2][4];
#pragma unroll
        for (int a = 0; a < PHO; ++a)
#pragma unroll
          for (int b = 0; b < 2; ++b)
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[a][b][k] = 0.f;
        const uint2 ysb = *reinterpret_cast<const uint2*>(rb + kOffYStart);
        auto byte_of = [](uint2 v, int i) -> int { return (int)(((i < 4 ? v.x : v.y) >> (8 * (i & 3))) & 255u); };
        const uint32_t* wy2 = reinterpret_cast<const uint32_t*>(rb + kOffWy);
        const uint32_t* xwtab = reinterpret_cast<const uint32_t*>(rb + kOffXw2) + lane;
        for (int tile = 0; tile < ntiles; ++tile) {
          const int il = min(tile * 8 + rr, nxs - 1);
          const int xl = rb[kOffXs + il];
          const uint32_t laddr = slice_u32 + xl * kPixBytes + (((mi ^ (xl >> 1)) & 3) << 4);
          const uint32_t amax = laddr + (H - 1) * rowbytes;
          uint32_t xw2u;
          if (tile < kTabTiles) {
            xw2u = xwtab[tile * 32];
          } else {
            const int pw = gq * STEP;
            const bool pw_ok = gq < PWO;
            xw2u = xw_pair(rb, tile * 8 + 2 * t, nxs, pw_ok ? rb[kOffXStart + pw] : 0, pw_ok ? rb[kOffXCount + pw] : 0,
                           rf + kOffWx / 4 + (pw_ok ? pw : 0) * kTaps);
          }
          const __nv_bfloat162 xw2 = *reinterpret_cast<const __nv_bfloat162*>(&xw2u);
          uint32_t addr[PHO];
#pragma unroll
          for (int pho = 0; pho < PHO; ++pho) addr[pho] = laddr + byte_of(ysb, pho * STEP) * rowbytes;
          // All PHO rows advance together over ky < max(ny): PHO independent LDSM -> HMMA chains per iteration and
          // no branch inside.  Rows whose window is shorter read a clamped (valid) map row against a zero weight
          // (the wy table is zero padded to kTaps).
          for (int ky = 0; ky < nymax; ++ky) {
#pragma unroll
            for (int pho = 0; pho < PHO; ++pho) {
              const uint32_t a2 = wy2[pho * STEP * kTaps + ky];
              const __nv_bfloat162 b2 = __hmul2(*reinterpret_cast<const __nv_bfloat162*>(&a2), xw2);
              const uint32_t b0 = *reinterpret_cast<const uint32_t*>(&b2);
              uint32_t r0, r1, r2, r3;
              asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                           : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr[pho]) : "memory");
              mma_bf16_1688(acc[pho][0], r0, r1, b0);
              mma_bf16_1688(acc[pho][1], r2, r3, b0);
              addr[pho] = min(addr[pho] + (uint32_t)rowbytes, amax);
            }
          }
        }
        // epilogue: D (ch16 x pw8 fp32) x2 -> bf16 -> stmatrix.trans -> [pw][32 ch] rows; all PHO rows are staged
        // first, then leave as back-to-back 16 B streaming stores
#pragma unroll
        for (int pho = 0; pho < PHO; ++pho) {
          const uint32_t m0 = pack_bf16x2(acc[pho][0][0], acc[pho][0][1]), m1 = pack_bf16x2(acc[pho][0][2], acc[pho][0][3]);
          const uint32_t m2 = pack_bf16x2(acc[pho][1][0], acc[pho][1][1]), m3 = pack_bf16x2(acc[pho][1][2], acc[pho][1][3]);
          asm volatile("stmatrix.sync.aligned.m8n8.x4.trans.shared.b16 [%0], {%1,%2,%3,%4};"
                       ::"r"(my_stage_u32 + pho * 512 + rr * 64 + (((mi ^ (rr >> 1)) & 3) << 4)), "r"(m0), "r"(m1), "r"(m2), "r"(m3)
                       : "memory");
        }
        __syncwarp();
        if (lane < PWO * 4) {
          const int slot = lane >> 2, o = lane & 3;
          const unsigned char* src = my_stage + slot * 64 + (((o ^ (slot >> 1)) & 3) << 4);
          __nv_bfloat16* dst = obase + (size_t)slot * C + o * 8;
#pragma unroll
          for (int pho = 0; pho < PHO; ++pho) {
            const uint4 v = *reinterpret_cast<const uint4*>(src + pho * 512);
            asm volatile("st.global.cs.v4.b32 [%0], {%1,%2,%3,%4};"
                         ::"l"(dst + (size_t)pho * PWO * C), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
          }
        }
      } else {
        // rare shapes (sparse fixed sampling grids, windows wider than the tables): per-sample path, lane <-> channel
        const RoiGeom g = roi_geom(rois + 5 * (size_t)roi, scale, sampling_ratio, aligned, PH, PW);
        const float inv = 1.0f / g.count;
        const int oct = lane >> 3, sub = (lane & 7) * 2;
        auto px = [&](int y, int x) -> float {
          const unsigned char* p = s_slice + y * rowbytes + x * kPixBytes + (((oct ^ (x >> 1)) & 3) << 4) + sub;
          return __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(p));
        };
        for (int pho = 0; pho < PHO; ++pho)
          for (int pwo = 0; pwo < PWO; ++pwo) {
            float acc = 0.f;
            for (int iy = 0; iy < g.gh; ++iy) {
              const AxisTap ty = make_tap(sample_coord(g.start_h, pho * STEP, g.bin_h, iy, g.gh), H, 1);
              for (int ix = 0; ix < g.gw; ++ix) {
                const AxisTap tx = make_tap(sample_coord(g.start_w, pwo * STEP, g.bin_w, ix, g.gw), W, 1);
                acc += ty.wlo * tx.wlo * px(ty.lo, tx.lo) + ty.wlo * tx.whi * px(ty.lo, tx.hi) +
                       ty.whi * tx.wlo * px(ty.hi, tx.lo) + ty.whi * tx.whi * px(ty.hi, tx.hi);
              }
            }
            obase[((size_t)pho * PWO + pwo) * C + lane] = __float2bfloat16_rn(acc * inv);
          }
      }
      __syncwarp();
      idx = nidx;
      nidx = nidx < sg.len ? __shfl_sync(0xffffffffu, nnidx, 0) : nidx;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int make_slice_map(CUtensorMap* m, const void* ptr, int N, int C, int H, int W, int Wp) {
  static PFN_encodeTiled enc = nullptr;
  if (!enc) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess) {
      set_error("roi_align_fwd: cuTensorMapEncodeTiled unavailable");
      return B200_ERR_CUDA;
    }
    enc = (PFN_encodeTiled)p;
  }
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)kSliceCh, (cuuint32_t)Wp, (cuuint32_t)H, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("roi_align_fwd: cuTensorMapEncodeTiled (slice) failed (%d)", (int)r);
    return B200_ERR_CUDA;
  }
  return B200_OK;
}

static size_t slice_smem_bytes(int H, int Wp) {
  return (size_t)H * Wp * kPixBytes + kSliceWarps * kRecRing * kRecBytes + kSliceWarps * kStageBytes + 1024;
}

int launch_roi_slice_prepare(const float* rois, unsigned char* recs, int R, int H, int W, int PH, int PW, int bin_step,
                             float scale, int sr, int aligned, cudaStream_t st) {
  roi_slice_prepare_kernel<<<ceil_div(R, 8), 256, 0, st>>>(rois, recs, R, H, W, PH, PW, bin_step, scale, sr, aligned);
  B200_CUDA_LAUNCH_CHECK("roi_slice_prepare");
  return B200_OK;
}

size_t roi_slice_workspace_bytes(int R) { return align_up((size_t)max(R, 1) * kRecBytes, 256); }

// true when the slice-resident kernel can run this problem (bf16, channels-last in/out, ROIs grouped by image)
bool roi_slice_eligible(int C, int H, int W, int PH, int PW, int bin_step, const void* feat) {
  if (!(PH == 7 && PW == 7 && (bin_step == 1 || bin_step == 2))) return false;
  if (C % kSliceCh != 0 || H > 256 || W > 248 || ((uintptr_t)feat & 15) != 0) return false;
  const int Wp = (W + 7) & ~7;
  return slice_smem_bytes(H, Wp) <= 227 * 1024;
}

int launch_roi_fwd_slice(const __nv_bfloat16* feat, const float* rois, const int32_t* roi_offsets, __nv_bfloat16* out,
                         int N, int C, int H, int W, int R, int PH, int PW, int bin_step, float scale, int sr,
                         int aligned, void* workspace, cudaStream_t st) {
  const int Wp = (W + 7) & ~7;
  CUtensorMap fmap;
  int rc = make_slice_map(&fmap, feat, N, C, H, W, Wp);
  if (rc != B200_OK) return rc;
  unsigned char* recs = (unsigned char*)workspace;
  rc = launch_roi_slice_prepare(rois, recs, R, H, W, PH, PW, bin_step, scale, sr, aligned, st);
  if (rc != B200_OK) return rc;
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    B200_CUDA_CALL(cudaGetDevice(&dev));
    B200_CUDA_CALL(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const size_t smem = slice_smem_bytes(H, Wp);
  const long long T = (long long)R * (C / kSliceCh);
  const int grid = (int)min((long long)num_sms, T);
  if (bin_step == 1) {
    auto k = roi_align_fwd_slice_kernel<7, 7, 1>;
    B200_CUDA_CALL(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, kSliceWarps * 32, smem, st>>>(fmap, recs, rois, roi_offsets, out, N, C, H, W, Wp, PH, PW, scale, sr, aligned);
  } else {
    auto k = roi_align_fwd_slice_kernel<4, 4, 2>;
    B200_CUDA_CALL(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, kSliceWarps * 32, smem, st>>>(fmap, recs, rois, roi_offsets, out, N, C, H, W, Wp, PH, PW, scale, sr, aligned);
  }
  B200_CUDA_LAUNCH_CHECK("roi_align_fwd_slice");
  return B200_OK;
}

}  // namespace b200
