// gt_resolve.cuh - do_map's coordinate handling (gstgeometrictransform.c:167-207) for kernels that evaluate maps in
// fp64 on the device (gt_device_maps.cu, diffuse.cu). TUs including this are compiled with --fmad=false.
#pragma once
#include <stdint.h>

namespace {

__device__ __forceinline__ double clampd (double x, double lo, double hi) { return (x > hi) ? hi : ((x < lo) ? lo : x); }   // CLAMP
// (int) of a double as the reference's x86-64 build evaluates it (cvttsd2si): NaN and out-of-range give INT_MIN
__device__ __forceinline__ int d2i (double x) { return (x > -2147483649.0 && x < 2147483648.0) ? (int) x : (int) 0x80000000; }
// geometricmath.c:171-180
__device__ __forceinline__ double mod_float (double a, double b) {
  int n = d2i (a / b);
  a -= n * b;
  if (a < 0) return a + b;
  return a;
}
// do_map's policy, truncation and bounds test (gstgeometrictransform.c:167-207), as gt_maps.cpp's resolve_one
__device__ __forceinline__ int32_t resolve_one (double ix, double iy, int width, int height, int off_edge) {
  if (off_edge == 1) {
    ix = clampd (ix, 0, width - 1);
    iy = clampd (iy, 0, height - 1);
  } else if (off_edge == 2) {
    ix = mod_float (ix, width);
    iy = mod_float (iy, height);
    if (ix < 0) ix += width;
    if (iy < 0) iy += height;
  }
  const int tx = d2i (ix), ty = d2i (iy);
  return (tx >= 0 && tx < width && ty >= 0 && ty < height) ? ty * width + tx : -1;
}

}  // namespace
