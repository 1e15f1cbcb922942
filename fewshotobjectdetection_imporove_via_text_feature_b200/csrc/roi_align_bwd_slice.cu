// P1b: ROIAlign backward, per-pixel CSR gather (bf16, channels-last) — atomic-free, deterministic.  The fine-tune path's
// default until the pixel-tile gather (roi_align_bwd_tile.cu) replaced it for C % 64 == 0; still "roi_align_bwd_impl" = 1
// and the path of every other channel count.  Reference: autograd of the roi_align call at defrcn/modeling/roi_heads/roi_heads.py:340
// (torchvision scatters with atomicAdd; fine-tuning reaches it with BACKWARD_SCALE = 0.001 through the GDL).
//
//   grad_feat[n,y,x,c] = sum over ROIs r of image n (index order), computed bins (ph,pw) whose pixel windows contain
//                        (y,x):  a_r,ph[y] * b_r,pw[x] * g[r,ph,pw,c]          (separable identity, roi_align_bwd.cu)
// Launches:
//   1. roi_slice_prepare_kernel (roi_align_slice.cu): the per-ROI geometry records the forward uses.
//   2. roi_bwd_csr_build_kernel<false> then <true>: CTA = (map row, image, one of kCsrGroups groups of the image's ROIs),
//      thread = pixel.  The ROIs of the group that touch the row are compacted in index order into shared memory
//      together with their row weights and horizontal tables, then every thread walks them from shared memory.  The
//      first launch counts, the second places each (pixel, group) by prefix sums of the counts and writes the pixel's
//      ordered list of (gradient row index, weight) pairs.  Geometry is channel independent: built once, not once
//      per channel chunk.
//   3. roi_bwd_csr_gather_kernel: warp = (pixel, 256 channels), lane = 8 channels.  Lanes fetch 32 list entries with
//      one coalesced load, broadcast them by shuffle and stream 16 B of gradient per entry into fp32 registers;
//      every lane is busy on every entry (the row-gather kernel this replaces idled ~3/4 of its lanes on ROIs that
//      did not cover their pixel: 0.99 ms on the bench shape, see profiles/).
// Each output element has exactly one writer and a fixed summation order: bitwise reproducible, nothing zero-filled
// or atomically accumulated in floating point (the only atomics are integer list-size totals).
#include "common.cuh"
#include "roi_geom.cuh"
#include "roi_slice_rec.cuh"

namespace b200 {

constexpr int kCsrMaxW = 256;         // map width handled by one CTA of the build kernel
constexpr int kCsrGroups = 8;         // ROI groups per image: CTAs of the list builder per map row
constexpr int kCsrSlots = 64;         // ROIs staged per round of a builder CTA
constexpr int kCsrLaneCh = 8;         // channels per lane of the gather kernel (one 16-byte load)
constexpr int kCsrWarps = 8;          // pixels per CTA of the gather kernel

struct CsrEntry {
  int row;        // (roi * PHO + pho) * PWO + pwo : row of the (strided) gradient tensor, C channels each
  float w;        // a_ph[y] / count * b_pw[x]
};

// Per-sample path for the rare shapes the tables do not cover (sparse fixed sampling grids, windows wider than kTaps):
// visits every computed bin of ROI r whose samples touch pixel (y, x) with non-zero weight.
template <typename F>
__device__ __forceinline__ void for_each_bin_sampled(const float* __restrict__ rois, int r, int y, int x, int H, int W, int PH,
                                                     int PW, int PHO, int PWO, int bin_step, float scale,
                                                     int sampling_ratio, int aligned, F&& emit) {
  const RoiGeom q = roi_geom(rois + 5 * (size_t)r, scale, sampling_ratio, aligned, PH, PW);
  const float inv = 1.0f / q.count;
  for (int pho = 0; pho < PHO; ++pho) {
    float wyv = 0.f;
    for (int iy = 0; iy < q.gh; ++iy) {
      const AxisTap ty = make_tap(sample_coord(q.start_h, pho * bin_step, q.bin_h, iy, q.gh), H, 1);
      wyv += (ty.lo == y ? ty.wlo : 0.f) + (ty.hi == y ? ty.whi : 0.f);
    }
    if (wyv == 0.f) continue;
    for (int pwo = 0; pwo < PWO; ++pwo) {
      float wxv = 0.f;
      for (int ix = 0; ix < q.gw; ++ix) {
        const AxisTap tx = make_tap(sample_coord(q.start_w, pwo * bin_step, q.bin_w, ix, q.gw), W, 1);
        wxv += (tx.lo == x ? tx.wlo : 0.f) + (tx.hi == x ? tx.whi : 0.f);
      }
      const float w = wyv * wxv * inv;
      if (w != 0.f) emit((r * PHO + pho) * PWO + pwo, w);
    }
  }
}

// grid (ceil(H / RB), N, kCsrGroups), block = RB x Wp threads (Wp = W rounded up to whole warps, RB rows so that the
// CTA has ~256 threads).  CTA = (block of map rows, image, group of the image's ROIs), thread = pixel.
// FILL = false: counts[group][pixel] and row_total[row] (integer atomics);  FILL = true: the ordered (gradient row,
// weight) lists, placed by prefix sums of those counts, and lists[pixel] = {offset, count}.
template <bool FILL>
__global__ void __launch_bounds__(kCsrMaxW)
roi_bwd_csr_build_kernel(const unsigned char* __restrict__ recs, const float* __restrict__ rois,
                         const int32_t* __restrict__ roi_offsets, int* __restrict__ counts, unsigned int* __restrict__ row_total,
                         int2* __restrict__ lists, CsrEntry* __restrict__ entries, unsigned int capacity, int N, int H, int W,
                         int Wp, int RB, int PH, int PW, int bin_step, float scale, int sampling_ratio, int aligned) {
  __shared__ __align__(16) unsigned char s_rec[kCsrSlots][kRecBwdBytes];   // leading part of the staged geometry records
  __shared__ int s_roi[kCsrSlots];
  __shared__ int s_warp[kCsrMaxW / 32];
  __shared__ unsigned int s_red[kCsrMaxW / 32];
  const int n = blockIdx.y, grp = blockIdx.z;
  const int yb = blockIdx.x * RB, ry = threadIdx.x / Wp, x = threadIdx.x - ry * Wp, y = yb + ry;
  const bool px_ok = y < H && x < W;
  const int r0 = roi_offsets[n], r1 = roi_offsets[n + 1];
  const int per = (r1 - r0 + kCsrGroups - 1) / kCsrGroups;
  const int gb = r0 + grp * per, ge = min(r1, gb + per);
  const int PHO = (PH + bin_step - 1) / bin_step, PWO = (PW + bin_step - 1) / bin_step;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5, wpr = Wp >> 5;
  const size_t npix = (size_t)N * H * W;
  const size_t pix = ((size_t)n * H + min(y, H - 1)) * W + min(x, W - 1);
  unsigned int pos = 0;
  int cnt = 0;

  if (FILL) {
    // offset of this (pixel, group): lists of earlier map rows + earlier pixels of the row + earlier groups
    unsigned int before_rows = 0;
    for (int i = threadIdx.x; i < n * H + yb; i += blockDim.x) before_rows += row_total[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) before_rows += __shfl_xor_sync(0xffffffffu, before_rows, o);
    if (lane == 0) s_red[warp] = before_rows;
    int tot = 0, mine_before = 0;
    if (px_ok)
      for (int g2 = 0; g2 < kCsrGroups; ++g2) {
        const int c = counts[(size_t)g2 * npix + pix];
        tot += c;
        if (g2 < grp) mine_before += c;
      }
    int incl = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    unsigned int base = 0;
    for (int i = 0; i < nwarps; ++i) base += s_red[i];
    for (int i = 0; i < ry && yb + i < H; ++i) base += row_total[n * H + yb + i];      // earlier rows of this CTA
    int before = 0;
    for (int i = ry * wpr; i < warp; ++i) before += s_warp[i];                           // earlier warps of this row
    const unsigned int list_begin = base + (unsigned)(before + incl - tot);
    pos = list_begin + (unsigned)mine_before;
    // a list that would run past the workspace is dropped whole; the capacity is the worst case, so this never fires
    if (grp == 0 && px_ok) lists[pix] = make_int2((int)list_begin, list_begin + (unsigned)tot <= capacity ? tot : 0);
  }

  for (int rb = gb; rb < ge; rb += kCsrSlots) {
    __syncthreads();                                   // the previous round's slots are no longer read
    // ---- ordered compaction of the ROIs of this round that touch the row block (first two warps, lane <-> ROI) -----
    int hit = 0;
    const int r = rb + threadIdx.x;
    if (threadIdx.x < kCsrSlots && r < ge) {
      const uint4 h = __ldg(reinterpret_cast<const uint4*>(recs + (size_t)r * kRecBytes + kOffYExt));   // yext, xext
      const int table = __ldg(reinterpret_cast<const int*>(recs + (size_t)r * kRecBytes + kOffFlags));
      hit = table ? ((int)h.x <= min(yb + RB, H) - 1 && (int)h.y >= yb && (int)h.z <= (int)h.w) : 1;
    }
    const unsigned int bal = __ballot_sync(0xffffffffu, hit);
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    const int nhit = s_warp[0] + s_warp[1];
    if (hit) s_roi[(warp ? s_warp[0] : 0) + __popc(bal & ((1u << lane) - 1u))] = r;
    __syncthreads();
    for (int i = threadIdx.x; i < nhit * (kRecBwdBytes / 16); i += blockDim.x) {
      const int k = i / (kRecBwdBytes / 16), j = i - k * (kRecBwdBytes / 16);
      reinterpret_cast<uint4*>(s_rec[k])[j] = __ldg(reinterpret_cast<const uint4*>(recs + (size_t)s_roi[k] * kRecBytes) + j);
    }
    __syncthreads();
    // ---- per-pixel walk: shared memory only on the table path ------------------------------------------------------
    if (px_ok) {
      for (int k = 0; k < nhit; ++k) {
        const unsigned char* rec = s_rec[k];
        const int roi = s_roi[k];
        if (*reinterpret_cast<const int*>(rec + kOffFlags)) {
          const int4 ext = *reinterpret_cast<const int4*>(rec + kOffYExt);
          if (y < ext.x || y > ext.y || x < ext.z || x > ext.w) continue;
          const uint32_t* wy2 = reinterpret_cast<const uint32_t*>(rec + kOffWy);
          const float* wx = reinterpret_cast<const float*>(rec + kOffWx);
          for (int pho = 0; pho < PHO; ++pho) {
            const int ph = pho * bin_step;
            const int ky = y - (int)rec[kOffYStart + ph];
            if ((unsigned)ky >= (unsigned)rec[kOffYCount + ph]) continue;
            const float a = __uint_as_float(wy2[ph * kTaps + ky] << 16);      // bf16(a / count), as the forward uses it
            if (a == 0.f) continue;
            for (int pwo = 0; pwo < PWO; ++pwo) {
              const int pw = pwo * bin_step;
              const int kx = x - (int)rec[kOffXStart + pw];
              if ((unsigned)kx >= (unsigned)rec[kOffXCount + pw]) continue;
              const float w = a * wx[pw * kTaps + kx];
              if (w == 0.f) continue;
              if (FILL) {
                if (pos < capacity) entries[pos] = CsrEntry{(roi * PHO + pho) * PWO + pwo, w};
                ++pos;
              } else {
                ++cnt;
              }
            }
          }
        } else {
          for_each_bin_sampled(rois, roi, y, x, H, W, PH, PW, PHO, PWO, bin_step, scale, sampling_ratio, aligned,
                               [&](int row, float w) {
                                 if (FILL) {
                                   if (pos < capacity) entries[pos] = CsrEntry{row, w};
                                   ++pos;
                                 } else {
                                   ++cnt;
                                 }
                               });
        }
      }
    }
  }
  if (!FILL) {
    if (px_ok) counts[(size_t)grp * npix + pix] = cnt;
    unsigned int tot = (unsigned)cnt;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
    if (lane == 0 && tot) atomicAdd(row_total + n * H + y, tot);       // a warp never spans two map rows
  }
}

// grid (ceil(N*H*W / kCsrWarps), ceil(C / 256)), block kCsrWarps warps
__global__ void __launch_bounds__(kCsrWarps * 32)
roi_bwd_csr_gather_kernel(const __nv_bfloat16* __restrict__ g, const int2* __restrict__ lists,
                          const CsrEntry* __restrict__ entries, __nv_bfloat16* __restrict__ grad_feat, int npix, int C) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pix = blockIdx.x * kCsrWarps + warp;
  const int c = (blockIdx.y * 32 + lane) * kCsrLaneCh;
  if (pix >= npix) return;
  const int2 lst = __ldg(lists + pix);
  const bool live = c < C;
  const __nv_bfloat16* gc = g + (live ? c : 0);
  // accumulators as four fp32 pairs: sm_100's packed FFMA2 (fma.rn.f32x2) does two channels per issue slot — the kernel is
  // issue-bound (ncu: 52 % issue, one unpack + one FMA per channel and entry before)
  unsigned long long acc2[4] = {0ull, 0ull, 0ull, 0ull};
  auto fma_entry = [&](const uint4 t, float w) {
    unsigned long long ww;
    asm("mov.b64 %0, {%1, %1};" : "=l"(ww) : "r"(__float_as_uint(w)));
    const uint32_t q[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      unsigned long long gg;
      asm("mov.b64 %0, {%1, %2};" : "=l"(gg) : "r"(q[i] << 16), "r"(q[i] & 0xffff0000u));
      asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc2[i]) : "l"(gg), "l"(ww));
    }
  };
  for (int e0 = 0; e0 < lst.y; e0 += 32) {
    const int m = min(32, lst.y - e0);
    int2 mine = make_int2(0, 0);
    if (lane < m) mine = __ldg(reinterpret_cast<const int2*>(entries) + lst.x + e0 + lane);
    int j = 0;
    for (; j + 4 <= m; j += 4) {                     // four independent 16-byte loads in flight per lane
      uint4 t[4];
      float w[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int row = __shfl_sync(0xffffffffu, mine.x, j + u);
        w[u] = __int_as_float(__shfl_sync(0xffffffffu, mine.y, j + u));
        t[u] = __ldg(reinterpret_cast<const uint4*>(gc + (size_t)row * C));
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) fma_entry(t[u], w[u]);
    }
    for (; j < m; ++j) {
      const int row = __shfl_sync(0xffffffffu, mine.x, j);
      const float w = __int_as_float(__shfl_sync(0xffffffffu, mine.y, j));
      fma_entry(__ldg(reinterpret_cast<const uint4*>(gc + (size_t)row * C)), w);
    }
  }
  float acc[8];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint32_t lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(acc2[i]));
    acc[2 * i] = __uint_as_float(lo);
    acc[2 * i + 1] = __uint_as_float(hi);
  }
  if (live) {
    uint4 o;
    __nv_bfloat162 h0 = __floats2bfloat162_rn(acc[0], acc[1]), h1 = __floats2bfloat162_rn(acc[2], acc[3]);
    __nv_bfloat162 h2 = __floats2bfloat162_rn(acc[4], acc[5]), h3 = __floats2bfloat162_rn(acc[6], acc[7]);
    o.x = *reinterpret_cast<uint32_t*>(&h0); o.y = *reinterpret_cast<uint32_t*>(&h1);
    o.z = *reinterpret_cast<uint32_t*>(&h2); o.w = *reinterpret_cast<uint32_t*>(&h3);
    *reinterpret_cast<uint4*>(grad_feat + (size_t)pix * C + c) = o;
  }
}

bool roi_bwd_slice_eligible(int C, int H, int W, int PH, int PW, int bin_step) {
  return PH <= 7 && PW <= 7 && bin_step >= 1 && C % 8 == 0 && H <= 256 && W <= kCsrMaxW;
}

// Worst case of the list entries: windows of consecutive bins overlap by at most 2 pixels per axis, so one ROI
// contributes at most (H + 2 PHO)(W + 2 PWO) (pixel, bin) pairs; ROIs on the table path (every ROI of a <= 56-pixel-high
// map region per bin, i.e. all of a 600 x 800 image) at most (PHO kTaps)(PWO kTaps).
static size_t csr_capacity(int R, int H, int W, int PHO, int PWO) {
  return (size_t)max(R, 1) * (size_t)(H + 2 * PHO) * (size_t)(W + 2 * PWO);
}

size_t roi_bwd_slice_workspace_bytes(int N, int H, int W, int R, int PH, int PW, int bin_step) {
  const int PHO = ceil_div(PH, bin_step), PWO = ceil_div(PW, bin_step);
  const size_t npix = (size_t)N * H * W;
  return align_up((size_t)max(R, 1) * kRecBytes, 256) + align_up(npix * sizeof(int2), 256) +
         align_up(npix * kCsrGroups * sizeof(int), 256) + align_up((size_t)N * H * sizeof(unsigned int), 256) +
         align_up(csr_capacity(R, H, W, PHO, PWO) * sizeof(CsrEntry), 256);
}

struct CsrPlan {
  unsigned char* recs;
  int2* lists;
  int* counts;
  unsigned int* row_total;
  CsrEntry* entries;
  size_t capacity;
};

static CsrPlan carve_plan(void* workspace, int N, int H, int W, int R, int PHO, int PWO) {
  const size_t npix = (size_t)N * H * W;
  unsigned char* p = (unsigned char*)workspace;
  CsrPlan pl;
  pl.recs = p;                          p += align_up((size_t)max(R, 1) * kRecBytes, 256);
  pl.lists = (int2*)p;                  p += align_up(npix * sizeof(int2), 256);
  pl.counts = (int*)p;                  p += align_up(npix * kCsrGroups * sizeof(int), 256);
  pl.row_total = (unsigned int*)p;      p += align_up((size_t)N * H * sizeof(unsigned int), 256);
  pl.entries = (CsrEntry*)p;
  pl.capacity = csr_capacity(R, H, W, PHO, PWO);
  return pl;
}

// geometry only (no gradient, no channels): may run ahead of the backward pass, e.g. on a side stream during the forward
int launch_roi_bwd_plan(const float* rois, const int32_t* roi_offsets, int N, int H, int W, int R, int PH, int PW,
                        int bin_step, float scale, int sr, int aligned, void* workspace, cudaStream_t st) {
  const int PHO = ceil_div(PH, bin_step), PWO = ceil_div(PW, bin_step);
  const CsrPlan pl = carve_plan(workspace, N, H, W, R, PHO, PWO);
  if (pl.capacity > 0x7fffffffull || (size_t)N * H * W > 0x7fffffffull) {
    set_error("roi_align_bwd: %d ROIs on a %d x %d map exceed the 2^31-entry list index", R, H, W);
    return B200_ERR_UNSUPPORTED;
  }
  int rc = launch_roi_slice_prepare(rois, pl.recs, R, H, W, PH, PW, bin_step, scale, sr, aligned, st);
  if (rc != B200_OK) return rc;
  B200_CUDA_CALL(cudaMemsetAsync(pl.row_total, 0, (size_t)N * H * sizeof(unsigned int), st));
  const int Wp = ceil_div(W, 32) * 32, RB = max(1, kCsrMaxW / Wp);
  const dim3 grid(ceil_div(H, RB), N, kCsrGroups);
  const int block = RB * Wp;
  roi_bwd_csr_build_kernel<false><<<grid, block, 0, st>>>(pl.recs, rois, roi_offsets, pl.counts, pl.row_total, pl.lists,
                                                          pl.entries, (unsigned int)pl.capacity, N, H, W, Wp, RB, PH, PW,
                                                          bin_step, scale, sr, aligned);
  B200_CUDA_LAUNCH_CHECK("roi_bwd_csr_count");
  roi_bwd_csr_build_kernel<true><<<grid, block, 0, st>>>(pl.recs, rois, roi_offsets, pl.counts, pl.row_total, pl.lists,
                                                         pl.entries, (unsigned int)pl.capacity, N, H, W, Wp, RB, PH, PW,
                                                         bin_step, scale, sr, aligned);
  B200_CUDA_LAUNCH_CHECK("roi_bwd_csr_fill");
  return B200_OK;
}

int launch_roi_bwd_gather(const __nv_bfloat16* g, const void* workspace, __nv_bfloat16* grad_feat, int N, int C, int H, int W,
                          int R, int PH, int PW, int bin_step, cudaStream_t st) {
  const int PHO = ceil_div(PH, bin_step), PWO = ceil_div(PW, bin_step);
  const CsrPlan pl = carve_plan(const_cast<void*>(workspace), N, H, W, R, PHO, PWO);
  const size_t npix = (size_t)N * H * W;
  roi_bwd_csr_gather_kernel<<<dim3(ceil_div((int)npix, kCsrWarps), ceil_div(C, 32 * kCsrLaneCh)), kCsrWarps * 32, 0, st>>>(
      g, pl.lists, pl.entries, grad_feat, (int)npix, C);
  B200_CUDA_LAUNCH_CHECK("roi_bwd_csr_gather");
  return B200_OK;
}

int launch_roi_bwd_slice(const __nv_bfloat16* g, const float* rois, const int32_t* roi_offsets, __nv_bfloat16* grad_feat,
                         int N, int C, int H, int W, int R, int PH, int PW, int bin_step, float scale, int sr, int aligned,
                         void* workspace, cudaStream_t st) {
  int rc = launch_roi_bwd_plan(rois, roi_offsets, N, H, W, R, PH, PW, bin_step, scale, sr, aligned, workspace, st);
  if (rc != B200_OK) return rc;
  return launch_roi_bwd_gather(g, workspace, grad_feat, N, C, H, W, R, PH, PW, bin_step, st);
}

}  // namespace b200
