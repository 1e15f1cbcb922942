// Per-ROI geometry record shared by the slice-resident ROIAlign kernels (forward: roi_align_slice.cu, backward:
// roi_align_bwd_slice.cu).  Built once per ROI by roi_slice_prepare_kernel.
#pragma once
#include "common.cuh"

namespace b200 {

constexpr int kSliceCh = 32;                 // channels per resident slice (64 B per pixel)
constexpr int kPixBytes = kSliceCh * 2;
constexpr int kRecBytes = 1024;              // per-ROI geometry record
constexpr int kRecChunks = kRecBytes / 16;
constexpr int kSliceWarps = 16;
constexpr int kRecRing = 2;                  // geometry records per warp (one in use, one in flight)
constexpr int kMaxXs = 64;                   // distinct pixel columns an ROI may touch on the table path
constexpr int kTaps = 9;                     // weight slots per bin and axis (sampling grid <= 8: any ROI of a 38x50 map)
constexpr int kTabTiles = 3;                 // 8-pixel tiles whose B-fragment weights are tabulated in the record
constexpr int kStageBytes = 7 * 512;         // epilogue staging per warp: PHO x [8 pw][32 ch] bf16

// record layout (bytes)
constexpr int kOffBatch = 0, kOffFlags = 4, kOffNxs = 8, kOffNyMax = 12;
constexpr int kOffYStart = 16, kOffYCount = 24, kOffXStart = 32, kOffXCount = 40, kOffXs = 48;
constexpr int kOffYExt = 112;                // int32 x2: first / last map row touched by the computed bins (last < first: none)
constexpr int kOffXExt = 120;                // int32 x2: first / last map column touched by the computed bins
constexpr int kRecBwdBytes = 640;            // leading part of the record the backward's list builder stages (header, Wy, Wx)
constexpr int kOffWy = 128;                  // u32 [7][kTaps], zero padded: bf16x2 (a,a), a = vertical weight / count
constexpr int kOffWx = 384;                  // fp32 [7][kTaps], zero padded: horizontal weights (tiles >= kTabTiles)
constexpr int kOffXw2 = 640;                 // u32 [kTabTiles][32 lanes]: bf16x2 B-fragment weights of lane (g,t)


// launches the prepare kernel: R records of kRecBytes into `recs`
int launch_roi_slice_prepare(const float* rois, unsigned char* recs, int R, int H, int W, int PH, int PW, int bin_step,
                             float scale, int sr, int aligned, cudaStream_t st);

}  // namespace b200
