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
//
// Maps that call libm (fisheye: pow; circle, rotate, twirl, kaleidoscope: atan2 / sin / cos; pinch: sin, pow; sphere:
// acos / asin / sin / tan; waterripple: sin) cannot be reproduced operation by operation - CUDA's libm and glibc's
// differ in the last ulps and do_map truncates. They are built on the GPU anyway, as a CERTIFIED table: the kernel
// evaluates the map once with CUDA's libm and once more per libm call with that call's result moved by 2^-40
// (~1400x the worst disagreement two < 2-ulp implementations can have), which gives a first-order bound on how far the
// host's coordinate can be from the GPU's. An entry whose coordinate is further than that bound (/64, + 1e-9 px) from
// the nearest integer, and away from the two places where a map is discontinuous in a libm result (circle's angle
// wrap, sphere's asin domain), truncates to the same index on both sides and is final. The few others (a few to a
// few thousand per 8K table: e.g. the centre row and column of fisheye, whose coordinates are exact integers) are
// appended to a list, evaluated by the host's own map function with glibc, and patched in. If more than 1/64 of the
// table is uncertain (rotate at angle 0: every coordinate sits within 1e-13 of an integer), the call reports
// B200VF_E_UNSUPPORTED and the element builds the table on the host as before. The result is entry-for-entry the
// host's table (tests/test_remap_gpu.py) at ~1/100 of its cost. `marble` needs no libm once the host has built its
// 16 KB of lattice and displacement tables, and joins the exact set.
//
#include "common.cuh"
#include "gt_resolve.cuh"
#include <string.h>
#include <math.h>
#include <vector>

namespace {

enum MapId { M_MIRROR, M_SQUARE, M_STRETCH, M_BULGE, M_TUNNEL, M_PERSPECTIVE, M_MARBLE,
  M_FISHEYE, M_CIRCLE, M_KALEIDOSCOPE, M_PINCH, M_ROTATE, M_SPHERE, M_TWIRL, M_WATERRIPPLE };
struct DevMap {
  int map, width, height, off_edge;
  double v[9];                          // element properties in the order of gt_maps.cpp's ElementDef::defaults
  double x_center, y_center, radius;    // GstCircleGeometricTransform (gstcirclegeometrictransform.c:177-192)
  double pcx, pcy, pr, pr2;             // its precalc (:144-157), evaluated on the host
  const double *tables;                 // marble: p[514], g2[514][2], sin[256], cos[256]
};

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
    case M_MARBLE: {                    // gstmarble.c:185-227 over geometricmath.c:95-165 (tables built by the host)
      const double *p = s.tables, *g2 = s.tables + 514, *sin_table = s.tables + 514 + 1028, *cos_table = sin_table + 256;
      const double xscale = s.v[0];
      const int N = 0x1000, BM = 0xff;
      double t = x / xscale + N;
      const int bx0 = d2i (t) & BM, bx1 = (bx0 + 1) & BM;
      const double rx0 = t - d2i (t), rx1 = rx0 - 1.0;
      t = y / xscale + N;
      const int by0 = d2i (t) & BM, by1 = (by0 + 1) & BM;
      const double ry0 = t - d2i (t), ry1 = ry0 - 1.0;
      const int i = (int) p[bx0], j = (int) p[bx1];
      const int b00 = (int) p[i + by0], b10 = (int) p[j + by0], b01 = (int) p[i + by1], b11 = (int) p[j + by1];
      const double sx = rx0 * rx0 * (3.0 - 2.0 * rx0), sy = ry0 * ry0 * (3.0 - 2.0 * ry0);
      double u = rx0 * g2[2 * b00] + ry0 * g2[2 * b00 + 1], v = rx1 * g2[2 * b10] + ry0 * g2[2 * b10 + 1];
      const double a = u + sx * (v - u);
      u = rx0 * g2[2 * b01] + ry1 * g2[2 * b01 + 1]; v = rx1 * g2[2 * b11] + ry1 * g2[2 * b11 + 1];
      const double b = u + sx * (v - u);
      int displacement = d2i (127 * (1 + 1.5 * (a + sy * (b - a))));
      displacement = displacement > 255 ? 255 : (displacement < 0 ? 0 : displacement);
      ix = x + sin_table[displacement];
      iy = y + cos_table[displacement];
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
  index[(size_t) y * s.width + x] = resolve_one (ix, iy, s.width, s.height, s.off_edge);
}

// ---------------------------------------------------------------- maps that call libm: certified evaluation
constexpr double kPi = 3.1415926535897932384626433832795028841971693993751;   // G_PI
__device__ __forceinline__ double triangle (double x) {      // geometricmath.c:182-190
  double r = mod_float (x, 1.0);
  return 2.0 * (r < 0.5 ? r : 1 - r);
}
// the result of libm call number i; evaluation k moves it by 2^-40 (k = -1: nobody moves)
__device__ __forceinline__ double lm (int i, int k, double v) { return i == k ? v * (1.0 + 0x1p-40) : v; }

struct MapOut {
  double ix, iy;
  int touched;      // bit 0 / 1: ix / iy depend on a libm result
  bool disc;        // sits on a discontinuity of the map in a libm result: the host decides
};

// One evaluation of map MAP at (x, y). The expressions are those of gt_maps.cpp (= the reference's *_map functions),
// every libm result passing through lm().
template <int MAP> __device__ __forceinline__ MapOut libm_map (const DevMap &s, int x, int y, int k)
{
  MapOut o; o.touched = 3; o.disc = false;
  if (MAP == M_FISHEYE) {               // gstfisheye.c:77-125 (pow is called twice on the same argument)
    double width = s.width, height = s.height;
    double nx = 2.0 * x / width - 1.0, ny = 2.0 * y / height - 1.0;
    double r = sqrt ((nx * nx + ny * ny) / 2.0);
    double p6 = lm (0, k, pow (r, 6.0));
    nx *= (0.33 + 0.1 * r * r + 0.57 * p6);
    ny *= (0.33 + 0.1 * r * r + 0.57 * p6);
    o.ix = 0.5 * (nx + 1.0) * width;
    o.iy = 0.5 * (ny + 1.0) * height;
  } else if (MAP == M_CIRCLE) {         // gstcircle.c:164-189
    double dx = x - s.pcx, dy = y - s.pcy;
    double distance = sqrt (dx * dx + dy * dy);
    double theta = lm (0, k, atan2 (-dy, -dx)) + s.v[0];
    theta = mod_float (theta, 2 * kPi);
    o.disc = !(theta > 1e-9 && theta < 2 * kPi - 1e-9);          // the wrap: 0 on one side, 2 pi on the other
    o.ix = s.width * theta / (s.v[1] + 0.0001);
    o.iy = s.height * (1 - (distance - s.pr) / ((int) s.v[2] + 0.0001));
    o.touched = 1;
  } else if (MAP == M_KALEIDOSCOPE) {   // gstkaleidoscope.c:165-195
    double angle = s.v[0], angle2 = s.v[1];
    int sides = (int) s.v[2];
    double dx = x - s.pcx, dy = y - s.pcy;
    double distance = sqrt (dx * dx + dy * dy);
    double theta = lm (0, k, atan2 (dy, dx)) - angle - angle2;
    theta = triangle (theta / kPi * sides * 0.5);
    if (s.pr != 0) {
      double radiusc = s.pr / lm (1, k, cos (theta));
      distance = radiusc * triangle (distance / radiusc);
    }
    theta += angle;
    o.ix = s.pcx + distance * lm (2, k, cos (theta));
    o.iy = s.pcy + distance * lm (3, k, sin (theta));
  } else if (MAP == M_PINCH) {          // gstpinch.c:136-174
    double dx = x - s.pcx, dy = y - s.pcy;
    double distance = dx * dx + dy * dy;
    if (distance > s.pr2 || distance == 0) { o.ix = x; o.iy = y; o.touched = 0; return o; }
    double d = sqrt (distance / s.pr2);
    double t = lm (1, k, pow (lm (0, k, sin (kPi * 0.5 * d)), -s.v[0]));
    dx *= t; dy *= t;
    o.ix = s.pcx + dx;
    o.iy = s.pcy + dy;
  } else if (MAP == M_ROTATE) {         // gstrotate.c:137-182
    double cox = 0.5 * s.width, coy = 0.5 * s.height;
    double xo = x - cox, yo = y - coy;
    double ao = lm (0, k, atan2 (yo, xo));
    double r = sqrt (xo * xo + yo * yo);
    double ai = ao + s.v[0];
    double xi = r * lm (1, k, cos (ai)), yi = r * lm (2, k, sin (ai));
    o.ix = xi + cox;
    o.iy = yi + coy;
  } else if (MAP == M_SPHERE) {         // gstsphere.c:137-186
    double dx = x - s.pcx, dy = y - s.pcy;
    double dx2 = dx * dx, dy2 = dy * dy;
    if (dy2 >= (s.pr2 - (s.pr2 * dx2) / s.pr2)) { o.ix = x; o.iy = y; o.touched = 0; return o; }
    double r_refraction = 1.0 / s.v[0];
    double z = sqrt ((1.0 - dx2 / s.pr2 - dy2 / s.pr2) * (s.pr2));
    double z2 = z * z;
    double angle = lm (0, k, acos (dx / sqrt (dx2 + z2)));
    double angle1 = kPi / 2 - angle;
    double arg = lm (1, k, sin (angle1)) * r_refraction;
    o.disc = o.disc || fabs (fabs (arg) - 1.0) < 1e-9;            // asin's domain edge: NaN on one side
    double angle2 = lm (2, k, asin (arg));
    angle2 = kPi / 2 - angle - angle2;
    o.ix = x - lm (3, k, tan (angle2)) * z;
    angle = lm (4, k, acos (dy / sqrt (dy2 + z2)));
    angle1 = kPi / 2 - angle;
    arg = lm (5, k, sin (angle1)) * r_refraction;
    o.disc = o.disc || fabs (fabs (arg) - 1.0) < 1e-9;
    angle2 = lm (6, k, asin (arg));
    angle2 = kPi / 2 - angle - angle2;
    o.iy = y - lm (7, k, tan (angle2)) * z;
  } else if (MAP == M_TWIRL) {          // gsttwirl.c:136-164
    double dx = x - s.pcx, dy = y - s.pcy;
    double distance = dx * dx + dy * dy;
    if (distance > s.pr2) { o.ix = x; o.iy = y; o.touched = 0; return o; }
    double d = sqrt (distance);
    double a = lm (0, k, atan2 (dy, dx)) + s.v[0] * (s.pr - d) / s.pr;
    o.ix = s.pcx + d * lm (1, k, cos (a));
    o.iy = s.pcy + d * lm (2, k, sin (a));
  } else {                              // gstwaterripple.c:162-195
    double dx = x - s.pcx, dy = y - s.pcy;
    double distance = dx * dx + dy * dy;
    if (distance > s.pr2) { o.ix = x; o.iy = y; o.touched = 0; return o; }
    double wavelength = s.v[2];
    double d = sqrt (distance);
    double amount = s.v[0] * lm (0, k, sin (d / wavelength * kPi * 2 - s.v[1]));
    amount *= (s.pr - d) / s.pr;
    if (d != 0) amount *= wavelength / d;
    o.ix = x + dx * amount;
    o.iy = y + dy * amount;
  }
  return o;
}

// Can the host's coordinate truncate differently? v: this side's value, bound: sum over the libm calls of what a 2^-40
// move of that call's result does to v, nn: how many of the evaluations gave NaN (of `evals`), n: frame extent.
__device__ __forceinline__ bool uncertain (double v, double bound, int nn, int evals, int n, int off_edge) {
  if (nn == evals) return false;                       // NaN whatever the last ulps say (x86: INT_MIN, unmapped)
  if (nn) return true;
  const double g = bound * (1.0 / 64) + 1e-9;
  if (!(g < 0.25) || !(fabs (v) < 1e15)) return true;  // too sensitive (a pole), or out of the range where fractions exist
  if (off_edge != 2 && (v < -1.0 - g || v > n + g)) return false;   // outside whatever the last ulps: unmapped, or clamped to the edge
  double f = v - floor (v);
  f = fmin (f, 1.0 - f);
  return f <= g;
}

template <int MAP, int NCALLS> __global__ void __launch_bounds__ (128)
gt_index_libm_kernel (const __grid_constant__ DevMap s, int32_t *__restrict__ index, unsigned int *__restrict__ n_uncertain,
    int32_t *__restrict__ uncertain_px, unsigned int cap)
{
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= s.width) return;
  const MapOut o = libm_map<MAP> (s, x, y, -1);
  bool ask_host = o.disc;
  if (o.touched) {
    double bx = 0, by = 0;
    int nnx = (o.ix != o.ix), nny = (o.iy != o.iy);
#pragma unroll 1
    for (int k = 0; k < NCALLS; k++) {
      const MapOut q = libm_map<MAP> (s, x, y, k);
      ask_host = ask_host || q.disc;
      nnx += (q.ix != q.ix); nny += (q.iy != q.iy);
      bx += fabs (q.ix - o.ix); by += fabs (q.iy - o.iy);
    }
    if (o.touched & 1) ask_host = ask_host || uncertain (o.ix, bx, nnx, NCALLS + 1, s.width, s.off_edge);
    if (o.touched & 2) ask_host = ask_host || uncertain (o.iy, by, nny, NCALLS + 1, s.height, s.off_edge);
  }
  const int px = y * s.width + x;
  if (ask_host) {
    const unsigned int slot = atomicAdd (n_uncertain, 1u);
    if (slot < cap) uncertain_px[slot] = px;
  }
  index[px] = resolve_one (o.ix, o.iy, s.width, s.height, s.off_edge);
}

__global__ void gt_index_patch_kernel (int32_t *__restrict__ index, const int32_t *__restrict__ px, const int32_t *__restrict__ value, unsigned int n) {
  const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) index[px[i]] = value[i];
}

struct DevMapDef { const char *name; int id; bool circle; int ncalls; int nprops; const char *props[9]; double defaults[9]; };
const DevMapDef kDevMaps[] = {          // ncalls: libm calls per evaluation (0: the exact set)
  { "mirror", M_MIRROR, false, 0, 1, { "mode" }, { 0.0 } },
  { "square", M_SQUARE, false, 0, 3, { "width", "height", "zoom" }, { 0.5, 0.5, 2.0 } },
  { "stretch", M_STRETCH, true, 0, 1, { "intensity" }, { 0.5 } },
  { "bulge", M_BULGE, true, 0, 1, { "zoom" }, { 3.0 } },
  { "tunnel", M_TUNNEL, true, 0, 0, { nullptr }, { 0 } },
  { "perspective", M_PERSPECTIVE, false, 0, 9, { "matrix-0", "matrix-1", "matrix-2", "matrix-3", "matrix-4", "matrix-5", "matrix-6",
      "matrix-7", "matrix-8" }, { 1, 0, 0, 0, 1, 0, 0, 0, 1 } },
  { "marble", M_MARBLE, false, 0, 4, { "x-scale", "y-scale", "amount", "turbulence" }, { 4.0, 4.0, 1.0, 1.0 } },
  { "fisheye", M_FISHEYE, false, 1, 0, { nullptr }, { 0 } },
  { "circle", M_CIRCLE, true, 1, 3, { "angle", "spread-angle", "height" }, { 0.0, 3.1415926535897932384626433832795028841971693993751, 20.0 } },
  { "kaleidoscope", M_KALEIDOSCOPE, true, 4, 3, { "angle", "angle2", "sides" }, { 0.0, 0.0, 3.0 } },
  { "pinch", M_PINCH, true, 2, 1, { "intensity" }, { 0.5 } },
  { "rotate", M_ROTATE, false, 3, 1, { "angle" }, { 0.0 } },
  { "sphere", M_SPHERE, true, 8, 1, { "refraction" }, { 1.5 } },
  { "twirl", M_TWIRL, true, 3, 1, { "angle" }, { 3.1415926535897932384626433832795028841971693993751 } },
  { "waterripple", M_WATERRIPPLE, true, 1, 3, { "amplitude", "phase", "wavelength" }, { 10.0, 0.0, 16.0 } },
};

template <int MAP, int NCALLS> void launch_libm (const DevMap &s, int32_t *d_index, unsigned int *d_count, int32_t *d_list,
    unsigned int cap, cudaStream_t st) {
  dim3 grid ((s.width + 127) / 128, s.height);
  gt_index_libm_kernel<MAP, NCALLS><<<grid, 128, 0, st>>> (s, d_index, d_count, d_list, cap);
}

thread_local long long g_last_uncertain = -1;

}  // namespace

B200VF_API int b200vf_gt_device_map_supported (const char *element) {
  if (!element) return 0;
  for (const auto &d : kDevMaps) if (!strcmp (d.name, element)) return d.ncalls ? 2 : 1;
  return 0;
}

B200VF_API long long b200vf_gt_device_last_uncertain (void) { return g_last_uncertain; }

B200VF_API int b200vf_gt_build_index_device (b200vf_ctx *ctx, const char *element, int width, int height,
    const char *const *prop_names, const double *prop_values, int nprops, int off_edge, int32_t *d_index, void *stream)
{
  B200VF_REQUIRE (ctx && element && d_index && width > 0 && height > 0 && nprops >= 0, B200VF_E_INVAL, "gt_build_index_device: bad argument");
  B200VF_REQUIRE (off_edge >= 0 && off_edge <= 2, B200VF_E_PROPERTY, "gt_build_index_device: off-edge-pixels %d", off_edge);
  B200VF_REQUIRE ((long long) width * height < 0x7fffffffll && height <= 65535, B200VF_E_INVAL, "gt_build_index_device: frame too large");
  const DevMapDef *def = nullptr;
  for (const auto &d : kDevMaps) if (!strcmp (d.name, element)) def = &d;
  B200VF_REQUIRE (def, B200VF_E_UNSUPPORTED, "gt_build_index_device: no device map for `%s`", element);
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
  // GstCircleGeometricTransform's precalc (gstcirclegeometrictransform.c:144-157), as MapState::circle_precalc
  s.pcx = s.x_center * width;
  s.pcy = s.y_center * height;
  s.pr = s.radius * 0.5 * sqrt ((double) (width * width + height * height));
  s.pr2 = s.pr * s.pr;
  cudaStream_t st = b200vf_stream (ctx, stream);
  g_last_uncertain = 0;
  if (!def->ncalls) {
    double *d_tables = nullptr;
    if (def->id == M_MARBLE) {
      std::vector<double> tables (2054);
      int rc = b200vf_gt_marble_tables (prop_names, prop_values, nprops, tables.data ());
      if (rc) return rc;
      B200VF_CHECK_CUDA (cudaMallocFromPoolAsync ((void **) &d_tables, tables.size () * sizeof (double), ctx->scratch_pool, st));
      // pageable source: the copy is staged before the call returns, `tables` may go out of scope
      B200VF_CHECK_CUDA (cudaMemcpyAsync (d_tables, tables.data (), tables.size () * sizeof (double), cudaMemcpyHostToDevice, st));
      s.tables = d_tables;
    }
    dim3 grid ((width + 255) / 256, height);
    gt_index_kernel<<<grid, 256, 0, st>>> (s, d_index);
    int rc = b200vf_launched (ctx, "gt_index_device");
    if (d_tables) cudaFreeAsync (d_tables, st);
    return rc;
  }

  // certified build: kernel, then the host's answer for the entries the kernel could not vouch for
  const size_t npx = (size_t) width * height;
  const unsigned int cap = (unsigned int) (npx / 64 > 4096 ? npx / 64 : 4096);
  unsigned int *d_count = nullptr;
  B200VF_CHECK_CUDA (cudaMallocFromPoolAsync ((void **) &d_count, (size_t) (1 + 2 * (size_t) cap) * 4, ctx->scratch_pool, st));
  int32_t *d_list = (int32_t *) (d_count + 1), *d_values = d_list + cap;
  struct Scratch { void *p; cudaStream_t st; ~Scratch () { cudaFreeAsync (p, st); } } scratch = { d_count, st };
  B200VF_CHECK_CUDA (cudaMemsetAsync (d_count, 0, 4, st));
  switch (def->id) {
    case M_FISHEYE: launch_libm<M_FISHEYE, 1> (s, d_index, d_count, d_list, cap, st); break;
    case M_CIRCLE: launch_libm<M_CIRCLE, 1> (s, d_index, d_count, d_list, cap, st); break;
    case M_KALEIDOSCOPE: launch_libm<M_KALEIDOSCOPE, 4> (s, d_index, d_count, d_list, cap, st); break;
    case M_PINCH: launch_libm<M_PINCH, 2> (s, d_index, d_count, d_list, cap, st); break;
    case M_ROTATE: launch_libm<M_ROTATE, 3> (s, d_index, d_count, d_list, cap, st); break;
    case M_SPHERE: launch_libm<M_SPHERE, 8> (s, d_index, d_count, d_list, cap, st); break;
    case M_TWIRL: launch_libm<M_TWIRL, 3> (s, d_index, d_count, d_list, cap, st); break;
    default: launch_libm<M_WATERRIPPLE, 1> (s, d_index, d_count, d_list, cap, st); break;
  }
  int rc = b200vf_launched (ctx, "gt_index_libm");
  if (rc) return rc;
  unsigned int count = 0;
  B200VF_CHECK_CUDA (cudaMemcpyAsync (&count, d_count, 4, cudaMemcpyDeviceToHost, st));
  B200VF_CHECK_CUDA (cudaStreamSynchronize (st));
  g_last_uncertain = count;
  if (count > cap) {
    b200vf_set_error ("gt_build_index_device: %u of %zu entries of `%s` sit within the libm error bound of an integer; "
        "build this table on the host (b200vf_gt_build_map)", count, npx, element);
    return B200VF_E_UNSUPPORTED;
  }
  if (!count) return B200VF_OK;
  std::vector<int32_t> px (count), values (count);
  B200VF_CHECK_CUDA (cudaMemcpyAsync (px.data (), d_list, (size_t) count * 4, cudaMemcpyDeviceToHost, st));
  B200VF_CHECK_CUDA (cudaStreamSynchronize (st));
  rc = b200vf_gt_host_index_at (element, width, height, prop_names, prop_values, nprops, off_edge, px.data (), count, values.data ());
  if (rc) return rc;
  B200VF_CHECK_CUDA (cudaMemcpyAsync (d_values, values.data (), (size_t) count * 4, cudaMemcpyHostToDevice, st));
  gt_index_patch_kernel<<<(count + 255) / 256, 256, 0, st>>> (d_index, d_list, d_values, count);
  rc = b200vf_launched (ctx, "gt_index_patch");
  if (rc) return rc;
  B200VF_CHECK_CUDA (cudaStreamSynchronize (st));
  return B200VF_OK;
}
