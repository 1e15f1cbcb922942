// P1b: ROIAlign backward, pixel-tile gather on the tensor cores (bf16 gradient, channels-last) — atomic-free and
// deterministic like the per-pixel CSR gather (roi_align_bwd_slice.cu), which it replaces as the fine-tune default.
// Reference: autograd of the roi_align call at defrcn/modeling/roi_heads/roi_heads.py:340 (torchvision scatters with
// atomicAdd; fine-tuning reaches it with BACKWARD_SCALE = 0.001 through the GDL).
//
// Why: the per-pixel gather reads every gradient row g[r,ph,pw,:] once per pixel of its bin window (~10 x), 1.3 GB of
// L2 -> SM traffic for 134 MB of gradient on the bench shape, and pays ~16 issue slots per (row, pixel, 8 channels).
// Here the unit is a 4 x 4-pixel tile of the map:
//     grad_feat[tile pixel m, c] = sum over the tile's entries e of  Wt[m, e] * g[row_e, c]
// where the entries are the (ROI, bin) pairs whose pixel window meets the tile (ROI index order, then bin order) and
// Wt[m, e] = a_{r,ph}[y_m] * b_{r,pw}[x_m] (/ count) is the separable bilinear weight (roi_align_bwd.cu), zero where the
// window misses the pixel.  That is a [16 x E] x [E x C] product per tile:
//   * every gradient row is read once per TILE its window meets (~2.3 x instead of ~10 x);
//   * the contraction runs as mma.sync m16n8k16 bf16 (fp32 accumulate): the 16 tile pixels are M, sixteen entries are K,
//     and a warp owns 64 channels.  The B operand needs no shared-memory transposition: lane (g, t) loads 16 bytes
//     (channels 8g..8g+7) of the rows of entries 2t, 2t+1, 2t+8, 2t+9 straight from global memory and one byte permute
//     per fragment register pairs channel 8g + j of two consecutive entries — the j-th of eight n-tiles takes channel
//     8g + j from lane group g.  The accumulators of lane (g, t) then are 16 consecutive channels of pixels g and g + 8:
//     two 16-byte stores each.  (A first version on m16n8k8 tf32 with fp32-exact operands was bound by the legacy tensor
//     pipe: 89 % hmma-active in ncu, 96 us; the bf16 shape does twice the entries per instruction);
//   * the weights are rounded to bf16 once (2^-9 relative, like the forward's own bf16 weight fragments), by the plan
//     builder, which stores them in fragment order (one uint4 per lane and 16 entries) next to the 16 row indices.
// Launches: roi_slice_prepare_kernel (the forward's per-ROI geometry records), roi_bwd_tile_build_kernel (CTA = tile:
// count, allocate the tile's blocks with ONE integer atomicAdd, ordered fill), roi_bwd_tile_gather_kernel.  The position
// of a tile's blocks in the plan buffer depends on the atomic's arrival order; their CONTENT, and so every bit of the
// result, does not.
#include "common.cuh"
#include "roi_geom.cuh"
#include "roi_slice_rec.cuh"

namespace b200 {

constexpr int kTileSide = 4;                  // tile = 4 x 4 map pixels = the M of one mma
constexpr int kBlkEntries = 16;               // entries per plan block = the K of one mma (bf16 m16n8k16)
constexpr int kBlkABytes = 32 * 16;           // uint4 per lane: A fragment {a0, a1, a2, a3}, bf16 pairs along K
constexpr int kBlkBytes = kBlkABytes + kBlkEntries * 4;   // + the gradient row of each entry
constexpr int kTileChunk = 64;                // channels per warp of the gather kernel
constexpr int kTileWarps = 8;
constexpr int kBuildThreads = 256;
constexpr int kOrderBuckets = 32;             // heaviest-first launch order of the gather: tiles bucketed by block count

__device__ __forceinline__ void mma_bf16_16816(float* d, const uint4 a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b0), "r"(b1));
}

// Visits, in (pho, pwo) order, every computed bin of ROI r whose pixel window meets the tile at (ty0, tx0).
// FILL = false: returns their number only.  FILL = true: emit(row, ay[4], bx[4], inv) with the per-axis weights of the
// tile's four rows / columns; the weight of tile pixel (i, j) is ay[i] * bx[j] * inv.
template <bool FILL, typename F>
__device__ __forceinline__ int tile_visit_roi(const unsigned char* __restrict__ recs, const float* __restrict__ rois, int r,
                                              int ty0, int tx0, int H, int W, int PH, int PW, int PHO, int PWO, int bin_step,
                                              float scale, int sampling_ratio, int aligned, F&& emit) {
  const unsigned char* rec = recs + (size_t)r * kRecBytes;
  const int table = __ldg(reinterpret_cast<const int*>(rec + kOffFlags));
  int n = 0;
  if (table) {
    const int4 ext = __ldg(reinterpret_cast<const int4*>(rec + kOffYExt));       // ylo, yhi, xlo, xhi of the computed bins
    if (ext.x > ty0 + kTileSide - 1 || ext.y < ty0 || ext.z > tx0 + kTileSide - 1 || ext.w < tx0) return 0;
    const uint4 hy = __ldg(reinterpret_cast<const uint4*>(rec + kOffYStart));    // ystart[8], ycount[8]
    const uint4 hx = __ldg(reinterpret_cast<const uint4*>(rec + kOffXStart));    // xstart[8], xcount[8]
    const unsigned long long ys8 = (unsigned long long)hy.x | ((unsigned long long)hy.y << 32);
    const unsigned long long yc8 = (unsigned long long)hy.z | ((unsigned long long)hy.w << 32);
    const unsigned long long xs8 = (unsigned long long)hx.x | ((unsigned long long)hx.y << 32);
    const unsigned long long xc8 = (unsigned long long)hx.z | ((unsigned long long)hx.w << 32);
    if (!FILL) {
      int ny = 0, nx = 0;
      for (int ph = 0; ph < PH; ph += bin_step) {
        const int s = (int)((ys8 >> (8 * ph)) & 0xffu), c = (int)((yc8 >> (8 * ph)) & 0xffu);
        ny += (c > 0 && s < ty0 + kTileSide && s + c > ty0) ? 1 : 0;
      }
      for (int pw = 0; pw < PW; pw += bin_step) {
        const int s = (int)((xs8 >> (8 * pw)) & 0xffu), c = (int)((xc8 >> (8 * pw)) & 0xffu);
        nx += (c > 0 && s < tx0 + kTileSide && s + c > tx0) ? 1 : 0;
      }
      return ny * nx;
    }
    const uint32_t* wy2 = reinterpret_cast<const uint32_t*>(rec + kOffWy);
    const float* wx = reinterpret_cast<const float*>(rec + kOffWx);
    for (int pho = 0; pho < PHO; ++pho) {
      const int ph = pho * bin_step;
      const int sy = (int)((ys8 >> (8 * ph)) & 0xffu), cy = (int)((yc8 >> (8 * ph)) & 0xffu);
      if (!(cy > 0 && sy < ty0 + kTileSide && sy + cy > ty0)) continue;
      float ay[kTileSide];
#pragma unroll
      for (int i = 0; i < kTileSide; ++i) {
        const int k = ty0 + i - sy;
        // bf16(a / count), exactly the vertical weight the forward multiplies with
        ay[i] = (unsigned)k < (unsigned)cy ? __uint_as_float(__ldg(wy2 + ph * kTaps + k) << 16) : 0.f;
      }
      for (int pwo = 0; pwo < PWO; ++pwo) {
        const int pw = pwo * bin_step;
        const int sx = (int)((xs8 >> (8 * pw)) & 0xffu), cx = (int)((xc8 >> (8 * pw)) & 0xffu);
        if (!(cx > 0 && sx < tx0 + kTileSide && sx + cx > tx0)) continue;
        float bx[kTileSide];
#pragma unroll
        for (int j = 0; j < kTileSide; ++j) {
          const int k = tx0 + j - sx;
          bx[j] = (unsigned)k < (unsigned)cx ? __ldg(wx + pw * kTaps + k) : 0.f;
        }
        emit((r * PHO + pho) * PWO + pwo, ay, bx, 1.0f);
        ++n;
      }
    }
    return n;
  }
  // per-sample path for the rare ROIs the tables do not cover (sparse fixed sampling grids, windows wider than kTaps)
  const RoiGeom q = roi_geom(rois + 5 * (size_t)r, scale, sampling_ratio, aligned, PH, PW);
  const float inv = 1.0f / q.count;
  for (int pho = 0; pho < PHO; ++pho) {
    float ay[kTileSide] = {0.f, 0.f, 0.f, 0.f};
    for (int iy = 0; iy < q.gh; ++iy) {
      const AxisTap t = make_tap(sample_coord(q.start_h, pho * bin_step, q.bin_h, iy, q.gh), H, 1);
#pragma unroll
      for (int i = 0; i < kTileSide; ++i) ay[i] += (t.lo == ty0 + i ? t.wlo : 0.f) + (t.hi == ty0 + i ? t.whi : 0.f);
    }
    if (ay[0] == 0.f && ay[1] == 0.f && ay[2] == 0.f && ay[3] == 0.f) continue;
    for (int pwo = 0; pwo < PWO; ++pwo) {
      float bx[kTileSide] = {0.f, 0.f, 0.f, 0.f};
      for (int ix = 0; ix < q.gw; ++ix) {
        const AxisTap t = make_tap(sample_coord(q.start_w, pwo * bin_step, q.bin_w, ix, q.gw), W, 1);
#pragma unroll
        for (int j = 0; j < kTileSide; ++j) bx[j] += (t.lo == tx0 + j ? t.wlo : 0.f) + (t.hi == tx0 + j ? t.whi : 0.f);
      }
      if (bx[0] == 0.f && bx[1] == 0.f && bx[2] == 0.f && bx[3] == 0.f) continue;
      if (FILL) emit((r * PHO + pho) * PWO + pwo, ay, bx, inv);
      ++n;
    }
  }
  return n;
}

// entry e of a tile whose blocks start at `blk0`: its gradient row, and its 16 pixel weights in fragment order — lane
// (g, t) holds a0 = W[pixel g][entries 2t, 2t+1], a1 = W[g + 8][2t, 2t+1], a2 = W[g][2t+8, 2t+9], a3 = W[g + 8][2t+8, 2t+9]
__device__ __forceinline__ void tile_write_entry(unsigned char* __restrict__ blocks, unsigned int blk0, int e, int row,
                                                 const float* ay, const float* bx, float inv) {
  unsigned char* blk = blocks + (size_t)(blk0 + (unsigned)(e / kBlkEntries)) * kBlkBytes;
  const int k = e % kBlkEntries, t = (k & 7) >> 1, hi = k >> 3;
  __nv_bfloat16* A = reinterpret_cast<__nv_bfloat16*>(blk);
  reinterpret_cast<int*>(blk + kBlkABytes)[k] = row;
#pragma unroll
  for (int m = 0; m < kTileSide * kTileSide; ++m)      // pixel m = (row m >> 2, column m & 3) of the tile
    A[(((m & 7) * 4 + t) * 4 + (m >> 3) + 2 * hi) * 2 + (k & 1)] = __float2bfloat16_rn(ay[m >> 2] * bx[m & 3] * inv);
}

// grid (tiles per image, N), block kBuildThreads; thread <-> ROI (strided over the image's ROIs)
__global__ void __launch_bounds__(kBuildThreads)
roi_bwd_tile_build_kernel(const unsigned char* __restrict__ recs, const float* __restrict__ rois,
                          const int32_t* __restrict__ roi_offsets, int2* __restrict__ tiles, unsigned int* __restrict__ counter,
                          int* __restrict__ order_lists, unsigned char* __restrict__ blocks, unsigned int capacity_blocks, int H, int W, int TXn, int PH,
                          int PW, int bin_step, float scale, int sampling_ratio, int aligned) {
  __shared__ int s_warp[kBuildThreads / 32];
  __shared__ unsigned int s_base;
  __shared__ int s_total;
  const int n = blockIdx.y, tile = blockIdx.x;
  const int ty0 = (tile / TXn) * kTileSide, tx0 = (tile % TXn) * kTileSide;
  const int r0 = roi_offsets[n], r1 = roi_offsets[n + 1];
  const int PHO = (PH + bin_step - 1) / bin_step, PWO = (PW + bin_step - 1) / bin_step;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  auto none = [](int, const float*, const float*, float) {};
  auto count_of = [&](int r) {
    return tile_visit_roi<false>(recs, rois, r, ty0, tx0, H, W, PH, PW, PHO, PWO, bin_step, scale, sampling_ratio, aligned, none);
  };

  // ---- pass 1: the tile's entry count, then its blocks ---------------------------------------------------------------
  int mine = 0;
  for (int r = r0 + (int)threadIdx.x; r < r1; r += kBuildThreads) mine += count_of(r);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
  if (lane == 0) s_warp[warp] = mine;
  __syncthreads();
  if (threadIdx.x == 0) {
    int total = 0;
    for (int i = 0; i < kBuildThreads / 32; ++i) total += s_warp[i];
    const unsigned int nblk = (unsigned)(total + kBlkEntries - 1) / kBlkEntries;
    unsigned int base = nblk ? atomicAdd(counter, nblk) : 0u;
    // the capacity is the worst case (launch function), so this never fires; a tile that would not fit is left empty
    const bool fits = base + nblk <= capacity_blocks;
    tiles[(size_t)n * gridDim.x + tile] = make_int2((int)base, fits ? (int)nblk : 0);
    // launch order of the gather (scheduling only, never the arithmetic): the tile joins the bucket of its block count;
    // counter[1 + bucket] = tiles in the bucket, order_lists[bucket][.] = their indices in arrival order
    const int bkt = min(kOrderBuckets - 1, (fits ? (int)nblk : 0) / 2);
    const int ntiles = (int)(gridDim.x * gridDim.y);
    order_lists[(size_t)bkt * ntiles + atomicAdd(counter + 1 + bkt, 1u)] = n * (int)gridDim.x + tile;
    s_base = base;
    s_total = fits ? total : 0;
  }
  __syncthreads();
  const int total = s_total;
  const unsigned int base = s_base;
  if (total == 0) return;

  // ---- pass 2: ordered fill (ROI index order: block-wide exclusive scan of the per-ROI counts, round by round) --------
  int running = 0;
  for (int rb = r0; rb < r1; rb += kBuildThreads) {
    const int r = rb + (int)threadIdx.x;
    const int c = r < r1 ? count_of(r) : 0;
    int incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    __syncthreads();                                   // s_warp of the previous round / of pass 1 is no longer read
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    int before = 0, round_total = 0;
#pragma unroll
    for (int i = 0; i < kBuildThreads / 32; ++i) {
      const int v = s_warp[i];
      before += i < warp ? v : 0;
      round_total += v;
    }
    if (c) {
      int e = running + before + incl - c;
      tile_visit_roi<true>(recs, rois, r, ty0, tx0, H, W, PH, PW, PHO, PWO, bin_step, scale, sampling_ratio, aligned,
                           [&](int row, const float* ay, const float* bx, float inv) {
                             tile_write_entry(blocks, base, e, row, ay, bx, inv);
                             ++e;
                           });
    }
    running += round_total;
  }
  // the unused slots of the last block: zero weights on a valid row
  const float zero[kTileSide] = {0.f, 0.f, 0.f, 0.f};
  const int padded = (total + kBlkEntries - 1) / kBlkEntries * kBlkEntries;
  for (int e = total + (int)threadIdx.x; e < padded; e += kBuildThreads) tile_write_entry(blocks, base, e, 0, zero, zero, 0.f);
}

// 16 bytes (NT = 8) or 8 bytes (NT = 4) of one gradient row
template <int NT> struct RowSeg;
template <> struct RowSeg<8> {
  uint32_t w[4];
  __device__ __forceinline__ void load(const __nv_bfloat16* p) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
    w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
  }
};
template <> struct RowSeg<4> {
  uint32_t w[2];
  __device__ __forceinline__ void load(const __nv_bfloat16* p) {
    const uint2 v = __ldg(reinterpret_cast<const uint2*>(p));
    w[0] = v.x; w[1] = v.y;
  }
};

// grid ceil(tiles * (C / (8 NT)) / kTileWarps), block kTileWarps warps; warp = (tile, 8 NT channels): NT n-tiles, lane
// (g, t) loads channels NT g .. NT g + NT - 1 of its four entries and feeds channel NT g + j to n-tile j, so that its
// accumulators are the 2 NT consecutive channels [2 NT t, 2 NT t + 2 NT) of tile pixels g and g + 8.
// kUnroll plan blocks per iteration, optionally double-buffered in registers, kMinCtas CTAs per SM.  Measured
// (profiles/r02_roi_bwd_tile_variants.txt): one block per iteration beats two (0.072 vs 0.094 ms), register double
// buffering and a per-warp cp.async ring in shared memory (2 x the L2 sectors: 16-byte LDGSTS requests do not coalesce
// into sectors across lanes; MIO throttle) were slower than the plain loop.  Launching the tiles heaviest first (the plan
// builder buckets them by block count) takes the tail off the last wave: 0.184 -> 0.159 ms for all 49 bins.
template <int kUnroll, bool kDouble, int kMinCtas, int NT>
__global__ void __launch_bounds__(kTileWarps * 32, kMinCtas)
roi_bwd_tile_gather_kernel(const __nv_bfloat16* __restrict__ g, const int2* __restrict__ tiles,
                           const unsigned char* __restrict__ blocks, __nv_bfloat16* __restrict__ grad_feat, int ntiles, int H,
                           int W, int TYn, int TXn, int C, const unsigned int* __restrict__ order_counts,
                           const int* __restrict__ order_lists) {
  constexpr int kChunk = 8 * NT;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int chunks = C / kChunk;
  const long long wi = (long long)blockIdx.x * kTileWarps + warp;
  if (wi >= (long long)ntiles * chunks) return;
  int tile = (int)(wi / chunks);
  const int chunk = (int)(wi - (long long)tile * chunks);
  if (order_counts != nullptr) {
    // heaviest tiles first (longest-processing-time order against the tail of the last wave): rank -> bucket -> tile
    const int rank = tile;
    const int c = (int)__ldg(order_counts + 1 + (kOrderBuckets - 1 - lane));      // lane 0 <-> the heaviest bucket
    int incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    const unsigned int m = __ballot_sync(0xffffffffu, incl > rank);               // the counts add up to ntiles > rank
    const int l = __ffs(m) - 1;
    const int before = __shfl_sync(0xffffffffu, incl - c, l);
    tile = __ldg(order_lists + (size_t)(kOrderBuckets - 1 - l) * ntiles + (rank - before));
  }
  const int gq = lane >> 2, t = lane & 3;
  const int2 td = __ldg(tiles + tile);
  const unsigned char* bp = blocks + (size_t)(unsigned)td.x * kBlkBytes;
  const int nblk = td.y;
  const __nv_bfloat16* gc = g + chunk * kChunk + gq * NT;

  float acc[NT][4];
#pragma unroll
  for (int j = 0; j < NT; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;

  // rows of the entries this lane loads, one iteration ahead of the gradient loads that need them
  int2 rlo[kUnroll], rhi[kUnroll];
  auto load_rows = [&](int b) {
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      rlo[u] = rhi[u] = make_int2(0, 0);
      if (b + u < nblk) {
        const int2* rows = reinterpret_cast<const int2*>(bp + (size_t)(b + u) * kBlkBytes + kBlkABytes);
        rlo[u] = __ldg(rows + t);          // entries 2t, 2t + 1
        rhi[u] = __ldg(rows + 4 + t);      // entries 2t + 8, 2t + 9
      }
    }
  };
  struct Buf {
    RowSeg<NT> d[kUnroll][4];
  };
  auto issue = [&](Buf& q, int b) {                    // rlo / rhi hold the rows of blocks [b, b + kUnroll)
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      if (b + u < nblk) {
        q.d[u][0].load(gc + (size_t)rlo[u].x * C);
        q.d[u][1].load(gc + (size_t)rlo[u].y * C);
        q.d[u][2].load(gc + (size_t)rhi[u].x * C);
        q.d[u][3].load(gc + (size_t)rhi[u].y * C);
      }
    }
  };
  auto compute = [&](const Buf& q, int b) {
    // the A fragments share their plan block (and cache lines) with the row indices fetched earlier
    uint4 af[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u)
      if (b + u < nblk) af[u] = __ldg(reinterpret_cast<const uint4*>(bp + (size_t)(b + u) * kBlkBytes) + lane);
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      if (b + u < nblk) {                               // warp-uniform
#pragma unroll
        for (int j = 0; j < NT; ++j) {
          // channel NT * gq + j of entries (2t, 2t + 1) and (2t + 8, 2t + 9): the lower half-word is the lower K index
          const uint32_t sel = (j & 1) ? 0x7632u : 0x5410u;
          const uint32_t b0 = __byte_perm(q.d[u][0].w[j >> 1], q.d[u][1].w[j >> 1], sel);
          const uint32_t b1 = __byte_perm(q.d[u][2].w[j >> 1], q.d[u][3].w[j >> 1], sel);
          mma_bf16_16816(acc[j], af[u], b0, b1);
        }
      }
    }
  };
  load_rows(0);
  if (kDouble) {
    Buf qa, qb;
    issue(qa, 0);
    load_rows(kUnroll);
    for (int b = 0; b < nblk; b += 2 * kUnroll) {
      issue(qb, b + kUnroll);
      load_rows(b + 2 * kUnroll);
      compute(qa, b);
      issue(qa, b + 2 * kUnroll);
      load_rows(b + 3 * kUnroll);
      compute(qb, b + kUnroll);
    }
  } else {
    Buf q;
    for (int b = 0; b < nblk; b += kUnroll) {
      issue(q, b);
      load_rows(b + kUnroll);
      compute(q, b);
    }
  }

  const int tpi = TYn * TXn;
  const int n = tile / tpi, tl = tile - n * tpi;
  const int ty0 = (tl / TXn) * kTileSide, tx0 = (tl % TXn) * kTileSide;
  const int x = tx0 + (gq & 3);
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int y = ty0 + (gq >> 2) + 2 * half;                 // tile pixel gq (c0, c1) / gq + 8 (c2, c3)
    if (y < H && x < W) {
      uint32_t p[NT];                                           // channels 2 NT t + [0, NT) from c0 / c2, + [NT, 2 NT) from c1 / c3
#pragma unroll
      for (int j = 0; j < NT; j += 2) {
        const __nv_bfloat162 lo = __floats2bfloat162_rn(acc[j][2 * half], acc[j + 1][2 * half]);
        const __nv_bfloat162 hi = __floats2bfloat162_rn(acc[j][2 * half + 1], acc[j + 1][2 * half + 1]);
        p[j >> 1] = *reinterpret_cast<const uint32_t*>(&lo);
        p[NT / 2 + (j >> 1)] = *reinterpret_cast<const uint32_t*>(&hi);
      }
      __nv_bfloat16* dst = grad_feat + ((size_t)(n * H + y) * W + x) * C + chunk * kChunk + 2 * NT * t;
      if (NT == 8) {
        reinterpret_cast<uint4*>(dst)[0] = make_uint4(p[0], p[1], p[2], p[3]);
        reinterpret_cast<uint4*>(dst)[1] = make_uint4(p[NT - 4], p[NT - 3], p[NT - 2], p[NT - 1]);
      } else {
        reinterpret_cast<uint4*>(dst)[0] = make_uint4(p[0], p[1], p[2], p[3]);
      }
    }
  }
}

// pipelining variant of the gather kernel ("roi_bwd_tile_variant" option: A/B measurement aid, every variant gives the
// same bits; profiles/r02_roi_bwd_tile_variants.txt)
int g_roi_bwd_tile_variant = 0;

bool roi_bwd_tile_eligible(int C, int H, int W, int PH, int PW, int bin_step) {
  return PH <= 7 && PW <= 7 && bin_step >= 1 && C % kTileChunk == 0 && H <= 256 && W <= 256;
}

struct TilePlan {
  unsigned char* recs;
  int2* tiles;
  unsigned int* counter;      // [0]: blocks allocated, [1 .. kOrderBuckets]: tiles per order bucket
  int* order_lists;           // [kOrderBuckets][ntiles]
  unsigned char* blocks;
  size_t capacity_blocks;
  int TYn, TXn;
};

// Worst case of the (tile, bin) pairs of one ROI: a bin window of L pixels meets at most (L + 6) / 4 tiles per axis, and
// the windows of one axis add up to at most size + 2 * bins + 2 pixels (consecutive windows overlap by <= 2 pixels).
static size_t tile_capacity_blocks(int N, int H, int W, int R, int PHO, int PWO) {
  const size_t per_roi = (size_t)ceil_div(H + 8 * PHO + 2, kTileSide) * (size_t)ceil_div(W + 8 * PWO + 2, kTileSide);
  const size_t ntiles = (size_t)N * ceil_div(H, kTileSide) * ceil_div(W, kTileSide);
  return ((size_t)max(R, 1) * per_roi + kBlkEntries - 1) / kBlkEntries + ntiles;      // + one partly filled block per tile
}

size_t roi_bwd_tile_workspace_bytes(int N, int H, int W, int R, int PH, int PW, int bin_step) {
  const int PHO = ceil_div(PH, bin_step), PWO = ceil_div(PW, bin_step);
  const size_t ntiles = (size_t)N * ceil_div(H, kTileSide) * ceil_div(W, kTileSide);
  return align_up((size_t)max(R, 1) * kRecBytes, 256) + align_up(ntiles * sizeof(int2), 256) + 256 +
         align_up(ntiles * kOrderBuckets * sizeof(int), 256) + tile_capacity_blocks(N, H, W, R, PHO, PWO) * kBlkBytes;
}

static TilePlan carve_tile_plan(void* workspace, int N, int H, int W, int R, int PHO, int PWO) {
  TilePlan pl;
  pl.TYn = ceil_div(H, kTileSide);
  pl.TXn = ceil_div(W, kTileSide);
  const size_t ntiles = (size_t)N * pl.TYn * pl.TXn;
  unsigned char* p = (unsigned char*)workspace;
  pl.recs = p;                       p += align_up((size_t)max(R, 1) * kRecBytes, 256);
  pl.tiles = (int2*)p;               p += align_up(ntiles * sizeof(int2), 256);
  pl.counter = (unsigned int*)p;     p += 256;
  pl.order_lists = (int*)p;          p += align_up(ntiles * kOrderBuckets * sizeof(int), 256);
  pl.blocks = p;
  pl.capacity_blocks = tile_capacity_blocks(N, H, W, R, PHO, PWO);
  return pl;
}

// geometry only (no gradient, no channels): may run ahead of the backward pass, e.g. on a side stream during the forward
int launch_roi_bwd_tile_plan(const float* rois, const int32_t* roi_offsets, int N, int H, int W, int R, int PH, int PW,
                             int bin_step, float scale, int sr, int aligned, void* workspace, cudaStream_t st) {
  const int PHO = ceil_div(PH, bin_step), PWO = ceil_div(PW, bin_step);
  const TilePlan pl = carve_tile_plan(workspace, N, H, W, R, PHO, PWO);
  if (pl.capacity_blocks > 0x7fffffffull || (size_t)R * PHO * PWO > 0x7fffffffull) {
    set_error("roi_align_bwd: %d ROIs on a %d x %d map exceed the 2^31-block plan index", R, H, W);
    return B200_ERR_UNSUPPORTED;
  }
  int rc = launch_roi_slice_prepare(rois, pl.recs, R, H, W, PH, PW, bin_step, scale, sr, aligned, st);
  if (rc != B200_OK) return rc;
  B200_CUDA_CALL(cudaMemsetAsync(pl.counter, 0, 256, st));
  roi_bwd_tile_build_kernel<<<dim3(pl.TYn * pl.TXn, N), kBuildThreads, 0, st>>>(
      pl.recs, rois, roi_offsets, pl.tiles, pl.counter, pl.order_lists, pl.blocks, (unsigned int)pl.capacity_blocks, H, W, pl.TXn, PH, PW,
      bin_step, scale, sr, aligned);
  B200_CUDA_LAUNCH_CHECK("roi_bwd_tile_build");
  return B200_OK;
}

int launch_roi_bwd_tile_gather(const __nv_bfloat16* g, const void* workspace, __nv_bfloat16* grad_feat, int N, int C, int H,
                               int W, int R, int PH, int PW, int bin_step, cudaStream_t st) {
  const int PHO = ceil_div(PH, bin_step), PWO = ceil_div(PW, bin_step);
  const TilePlan pl = carve_tile_plan(const_cast<void*>(workspace), N, H, W, R, PHO, PWO);
  const int ntiles = N * pl.TYn * pl.TXn;
#define B200_TILE_GATHER(U, DB, MINB, NT, ORDERED)                                                                   \
  do {                                                                                                                 \
    const long long warps = (long long)ntiles * (C / (8 * NT));                                                        \
    roi_bwd_tile_gather_kernel<U, DB, MINB, NT><<<(unsigned)((warps + kTileWarps - 1) / kTileWarps), kTileWarps * 32, 0, st>>>( \
        g, pl.tiles, pl.blocks, grad_feat, ntiles, H, W, pl.TYn, pl.TXn, C, (ORDERED) ? pl.counter : nullptr,            \
        pl.order_lists);                                                                                               \
  } while (0)
  switch (g_roi_bwd_tile_variant) {
    case 1: B200_TILE_GATHER(2, false, 2, 8, false); break;    // two plan blocks per iteration, 16 warps / SM: 0.094 ms (16 bins)
    case 2: B200_TILE_GATHER(1, false, 4, 4, false); break;    // 32-channel warps, 32 warps / SM: 0.083 ms
    case 3: B200_TILE_GATHER(1, false, 3, 8, false); break;    // the default's kernel, tiles in map order: 0.072 ms
    default: B200_TILE_GATHER(1, false, 3, 8, true); break;    // one block per iteration, 24 warps / SM, heaviest tiles first: 0.067 ms
  }
#undef B200_TILE_GATHER
  B200_CUDA_LAUNCH_CHECK("roi_bwd_tile_gather");
  return B200_OK;
}

}  // namespace b200
