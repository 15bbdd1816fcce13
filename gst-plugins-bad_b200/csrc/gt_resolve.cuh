// gt_resolve.cuh - do_map's coordinate handling (gstgeometrictransform.c:167-207) for kernels that evaluate maps in
// fp64 on the device (gt_device_maps.cu, diffuse.cu). TUs including this are compiled with --fmad=false.
#pragma once
#include <stdint.h>

namespace {

__device__ __forceinline__ double clampd (double x, double lo, double hi) { return (x > hi) ? hi : ((x < lo) ? lo : x); }   // CLAMP
// (int) of a double as the reference's x86-64 build evaluates it (cvttsd2si): truncation toward zero, NaN and
// out-of-range give INT_MIN. Done on the fp64 adder instead of F2I: the conversion unit issues one warp instruction per
// 8 cycles and was the limiter of diffuse (ncu: xu pipe saturated). x + (2^52 + 2^51) rounds x to the nearest integer
// and leaves it, two's complement, in the low word; one compare turns nearest into toward-zero.
__device__ __forceinline__ int d2i (double x) {
  if (!(fabs (x) < 2147483648.0)) return (int) 0x80000000;      // NaN too; (-2^31 - 1, -2^31] truncates to INT_MIN anyway
  const double M = 6755399441055744.0;
  const double t = x + M;
  unsigned int r = (unsigned int) __double2loint (t);            // unsigned: 2^31 - 0.5 rounds to 2^31 first, then steps back
  const double rn = t - M;
  if (x >= 0) { if (rn > x) r--; } else { if (rn < x) r++; }
  return (int) r;
}
// geometricmath.c:171-180
__device__ __forceinline__ double mod_float (double a, double b) {
  int n = d2i (a / b);
  a -= n * b;
  if (a < 0) return a + b;
  return a;
}
// do_map's policy, truncation and bounds test (gstgeometrictransform.c:167-207), as gt_maps.cpp's resolve_one:
// the source pixel (tx, ty), or false when the output keeps the cleared frame's value
__device__ __forceinline__ bool resolve_xy (double ix, double iy, int width, int height, int off_edge, int &tx, int &ty) {
  if (off_edge == 1) {
    ix = clampd (ix, 0, width - 1);
    iy = clampd (iy, 0, height - 1);
  } else if (off_edge == 2) {
    ix = mod_float (ix, width);
    iy = mod_float (iy, height);
    if (ix < 0) ix += width;
    if (iy < 0) iy += height;
  }
  tx = d2i (ix); ty = d2i (iy);
  return tx >= 0 && tx < width && ty >= 0 && ty < height;
}
__device__ __forceinline__ int32_t resolve_one (double ix, double iy, int width, int height, int off_edge) {
  int tx, ty;
  return resolve_xy (ix, iy, width, height, off_edge, tx, ty) ? ty * width + tx : -1;
}

}  // namespace
