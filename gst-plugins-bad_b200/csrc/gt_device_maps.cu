// gt_device_maps.cu - geometrictransform maps evaluated on the GPU.
//
// The reference builds its gdouble map on one CPU thread (gst_geometric_transform_generate_map,
// gstgeometrictransform.c:80-128: 3.6 s at 8K) every time a property changes (needs_remap), and the host path here
// (gt_maps.cpp) still pays 0.2 s + a 132 MB upload per change. For the maps whose arithmetic is +, -, *, /, sqrt,
// comparisons and double -> int truncation only - mirror, square, stretch, bulge, tunnel, perspective - the GPU
// evaluates the SAME expressions in IEEE fp64 (this file is compiled with --fmad=false: a contracted multiply-add
// would round differently; division and sqrt are correctly rounded on both sides) and writes the int32 gather table
// straight into HBM: ~0.1 ms at 8K, nothing crosses PCIe, a GstController animating a property costs a table rebuild
// per frame instead of a pipeline stall. tests/test_remap_gpu.py checks index-for-index equality with the host
// tables (which are bit-equal to the reference's maps) for every property set, policy and size tested there.
// Maps that call libm (pow, atan2, sin, cos: fisheye, circle, kaleidoscope, pinch, rotate, sphere, twirl,
// waterripple) stay on the host: glibc's results are not reproducible operation by operation.
//
// The gather itself stays table-driven (remap.cu): re-evaluating ~100 fp64 instructions per pixel per frame would
// cost more than reading 4 B/px (measured: profiles/r02_remap.md).
#include "common.cuh"
#include <string.h>

namespace {

enum MapId { M_MIRROR, M_SQUARE, M_STRETCH, M_BULGE, M_TUNNEL, M_PERSPECTIVE };
struct DevMap {
  int map, width, height, off_edge;
  double v[9];                          // element properties in the order of gt_maps.cpp's ElementDef::defaults
  double x_center, y_center, radius;    // GstCircleGeometricTransform (gstcirclegeometrictransform.c:177-192)
};

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
__device__ __forceinline__ double smoothstep (double e0, double e1, double x) {
  double t = clampd ((x - e0) / (e1 - e0), 0.0, 1.0);
  return t * t * (3.0 - 2.0 * t);
}

__global__ void __launch_bounds__ (256)
gt_index_kernel (const __grid_constant__ DevMap s, int32_t *__restrict__ index)
{
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= s.width) return;
  double ix, iy;
  switch (s.map) {
    case M_MIRROR: {                    // gstmirror.c:158-203
      double hw = s.width / 2.0 - 1.0, hh = s.height / 2.0 - 1.0;
      switch ((int) s.v[0]) {
        case 0: ix = (x > hw) ? s.width - 1.0 - x : x; iy = y; break;
        case 1: ix = (x > hw) ? x : s.width - 1.0 - x; iy = y; break;
        case 2: iy = (y > hh) ? s.height - 1.0 - y : y; ix = x; break;
        default: iy = (y > hh) ? y : s.height - 1.0 - y; ix = x; break;
      }
      break;
    }
    case M_SQUARE: {                    // gstsquare.c:158-192
      double width = s.width, height = s.height;
      double sw = s.v[0], sh = s.v[1], zoom = s.v[2];
      double nx = 2.0 * x / width - 1.0, ny = 2.0 * y / height - 1.0;
      double ax = nx < 0 ? -nx : nx, ay = ny < 0 ? -ny : ny;
      nx *= (1.0 / zoom) * (1.0 + (zoom - 1.0) * smoothstep (sw - 0.125, sw + 0.125, ax));
      ny *= (1.0 / zoom) * (1.0 + (zoom - 1.0) * smoothstep (sh - 0.125, sh + 0.125, ay));
      ix = 0.5 * (nx + 1.0) * width;
      iy = 0.5 * (ny + 1.0) * height;
      break;
    }
    case M_STRETCH: {                   // gststretch.c:132-176
      double width = s.width, height = s.height;
      double nx = 2.0 * (x / width - s.x_center), ny = 2.0 * (y / height - s.y_center);
      double r = sqrt (0.5 * (nx * nx + ny * ny));
      double a = 1.0 + (3.0 - 1.0) * s.v[0];
      double b = a - 1.0;
      nx *= a - b * smoothstep (0.0, s.radius, r);
      ny *= a - b * smoothstep (0.0, s.radius, r);
      ix = (0.5 * nx + s.x_center) * width;
      iy = (0.5 * ny + s.y_center) * height;
      break;
    }
    case M_BULGE: {                     // gstbulge.c:132-175
      double width = s.width, height = s.height, zoom = s.v[0];
      double nx = 2.0 * (x / width - s.x_center), ny = 2.0 * (y / height - s.y_center);
      double r = sqrt (0.5 * (nx * nx + ny * ny));
      double scale = 1.0 / (zoom + ((1.0 - zoom) * smoothstep (0, s.radius, r)));
      nx *= scale; ny *= scale;
      ix = (0.5 * nx + s.x_center) * width;
      iy = (0.5 * ny + s.y_center) * height;
      break;
    }
    case M_TUNNEL: {                    // gsttunnel.c:79-115 (the centre pixel: 0/0 = NaN, stays unmapped)
      double width = s.width, height = s.height;
      double m = width > height ? width : height;
      double nx = 2.0 * (x - s.x_center * width) / m, ny = 2.0 * (y - s.y_center * height) / m;
      double r = sqrt (0.5 * (nx * nx + ny * ny));
      nx *= clampd (r, 0.0, s.radius) / r;
      ny *= clampd (r, 0.0, s.radius) / r;
      ix = 0.5 * (nx) * m + s.x_center * width;
      iy = 0.5 * (ny) * m + s.y_center * height;
      break;
    }
    default: {                          // gstperspective.c:184-210
      double xp = (s.v[0] * x + s.v[1] * y + s.v[2]);
      double yp = (s.v[3] * x + s.v[4] * y + s.v[5]);
      double w = (s.v[6] * x + s.v[7] * y + s.v[8]);
      ix = xp / w;
      iy = yp / w;
      break;
    }
  }
  // do_map's policy and truncation (gstgeometrictransform.c:167-207), as b200vf_gt_resolve_map
  const int width = s.width, height = s.height;
  if (s.off_edge == 1) {
    ix = clampd (ix, 0, width - 1);
    iy = clampd (iy, 0, height - 1);
  } else if (s.off_edge == 2) {
    ix = mod_float (ix, width);
    iy = mod_float (iy, height);
    if (ix < 0) ix += width;
    if (iy < 0) iy += height;
  }
  const int tx = d2i (ix), ty = d2i (iy);
  index[(size_t) y * width + x] = (tx >= 0 && tx < width && ty >= 0 && ty < height) ? ty * width + tx : -1;
}

struct DevMapDef { const char *name; int id; bool circle; int nprops; const char *props[9]; double defaults[9]; };
const DevMapDef kDevMaps[] = {
  { "mirror", M_MIRROR, false, 1, { "mode" }, { 0.0 } },
  { "square", M_SQUARE, false, 3, { "width", "height", "zoom" }, { 0.5, 0.5, 2.0 } },
  { "stretch", M_STRETCH, true, 1, { "intensity" }, { 0.5 } },
  { "bulge", M_BULGE, true, 1, { "zoom" }, { 3.0 } },
  { "tunnel", M_TUNNEL, true, 0, { nullptr }, { 0 } },
  { "perspective", M_PERSPECTIVE, false, 9, { "matrix-0", "matrix-1", "matrix-2", "matrix-3", "matrix-4", "matrix-5", "matrix-6",
      "matrix-7", "matrix-8" }, { 1, 0, 0, 0, 1, 0, 0, 0, 1 } },
};

}  // namespace

B200VF_API int b200vf_gt_device_map_supported (const char *element) {
  if (!element) return 0;
  for (const auto &d : kDevMaps) if (!strcmp (d.name, element)) return 1;
  return 0;
}

B200VF_API int b200vf_gt_build_index_device (b200vf_ctx *ctx, const char *element, int width, int height,
    const char *const *prop_names, const double *prop_values, int nprops, int off_edge, int32_t *d_index, void *stream)
{
  B200VF_REQUIRE (ctx && element && d_index && width > 0 && height > 0 && nprops >= 0, B200VF_E_INVAL, "gt_build_index_device: bad argument");
  B200VF_REQUIRE (off_edge >= 0 && off_edge <= 2, B200VF_E_PROPERTY, "gt_build_index_device: off-edge-pixels %d", off_edge);
  B200VF_REQUIRE ((long long) width * height < 0x7fffffffll && height <= 65535, B200VF_E_INVAL, "gt_build_index_device: frame too large");
  const DevMapDef *def = nullptr;
  for (const auto &d : kDevMaps) if (!strcmp (d.name, element)) def = &d;
  B200VF_REQUIRE (def, B200VF_E_UNSUPPORTED, "gt_build_index_device: `%s` calls libm and is built on the host (b200vf_gt_build_map)", element);
  DevMap s;
  memset (&s, 0, sizeof s);
  s.map = def->id; s.width = width; s.height = height; s.off_edge = off_edge;
  s.x_center = 0.5; s.y_center = 0.5; s.radius = 0.35;
  for (int i = 0; i < def->nprops; i++) s.v[i] = def->defaults[i];
  for (int i = 0; i < nprops; i++) {
    B200VF_REQUIRE (prop_names && prop_values && prop_names[i], B200VF_E_INVAL, "gt_build_index_device: NULL property");
    const char *n = prop_names[i];
    if (!strcmp (n, "off-edge-pixels")) continue;
    bool found = false;
    for (int k = 0; k < def->nprops; k++) if (!strcmp (def->props[k], n)) { s.v[k] = prop_values[i]; found = true; }
    if (def->circle) {
      if (!strcmp (n, "x-center")) { s.x_center = prop_values[i]; found = true; }
      else if (!strcmp (n, "y-center")) { s.y_center = prop_values[i]; found = true; }
      else if (!strcmp (n, "radius")) { s.radius = prop_values[i]; found = true; }
    }
    B200VF_REQUIRE (found, B200VF_E_PROPERTY, "gt_build_index_device: element `%s` has no property `%s`", element, n);
  }
  cudaStream_t st = b200vf_stream (ctx, stream);
  dim3 grid ((width + 255) / 256, height);
  gt_index_kernel<<<grid, 256, 0, st>>> (s, d_index);
  return b200vf_launched (ctx, "gt_index_device");
}
