// ROIAlign geometry shared by the forward kernels (torchvision roi_align semantics, op-for-op rounding).
#pragma once
#include "common.cuh"

namespace b200 {

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

}  // namespace b200
