// gt_maps.cpp - host side of geometrictransform: the inverse-mapping functions and
// the resolution of the double (x,y) map into a compact int32 gather table.
//
// Why on the host: the reference builds this map once per caps/property change with
// glibc libm in double (gst_geometric_transform_generate_map,
// gst/geometrictransform/gstgeometrictransform.c:80-128) and truncates the doubles to
// pixel indices per frame (do_map, :167-207). Truncation makes the result sensitive to
// the last ulp of sqrt/pow/sin/atan2, so bit-exact parity needs the same libm
// (SURVEY.md D4, §8c-ii). The per-frame work - the gather - is the GPU kernel (remap.cu).
//
// Every map function below restates the expression order of the reference's
// `*_map` it cites; all arithmetic is IEEE double, no contraction (-ffp-contract=off).
#include "../csrc/common.cuh"
#include <math.h>
#include <string.h>
#include <stdlib.h>
#include <map>
#include <string>
#include <vector>
#include <thread>

namespace {

const double kPi = 3.1415926535897932384626433832795028841971693993751;   // G_PI

inline double clampd (double x, double lo, double hi) { return (x > hi) ? hi : ((x < lo) ? lo : x); }   // CLAMP

// geometricmath.c:171-201
double mod_float (double a, double b) {
  int n = (int) (a / b);
  a -= n * b;
  if (a < 0) return a + b;
  return a;
}
double triangle (double x) {
  double r = mod_float (x, 1.0);
  return 2.0 * (r < 0.5 ? r : 1 - r);
}
double smoothstep (double e0, double e1, double x) {
  double t = clampd ((x - e0) / (e1 - e0), 0.0, 1.0);
  return t * t * (3.0 - 2.0 * t);
}

// Perlin-style noise for `marble` (geometricmath.c:54-165). The reference seeds it from
// g_random_int (inherently unpinned); here a fixed-seed LCG, so frames are
// reproducible run to run but NOT comparable with any particular reference run.
struct Noise {
  double p[2 * 256 + 2];
  double g2[2 * 256 + 2][2];
  explicit Noise (uint32_t seed) {
    uint32_t s = seed;
    auto rnd = [&s] () { s = s * 1664525u + 1013904223u; return s >> 8; };
    const int B = 256;
    for (int i = 0; i < B; i++) {
      p[i] = i;
      for (int j = 0; j < 2; j++) g2[i][j] = ((int) (rnd () % (2 * B)) - B) / (double) B;
      double n = sqrt (g2[i][0] * g2[i][0] + g2[i][1] * g2[i][1]);
      if (n == 0) { g2[i][0] = 1.0; n = 1.0; }
      g2[i][0] /= n; g2[i][1] /= n;
    }
    for (int i = B - 1; i >= 0; i--) { int j = rnd () % B; double k = p[i]; p[i] = p[j]; p[j] = k; }
    for (int i = 0; i < B + 2; i++) { p[B + i] = p[i]; g2[B + i][0] = g2[i][0]; g2[B + i][1] = g2[i][1]; }
  }
  double at (double x, double y) const {
    const int N = 0x1000, BM = 0xff;
    double t = x + N;
    int bx0 = ((int) t) & BM, bx1 = (bx0 + 1) & BM;
    double rx0 = t - (int) t, rx1 = rx0 - 1.0;
    t = y + N;
    int by0 = ((int) t) & BM, by1 = (by0 + 1) & BM;
    double ry0 = t - (int) t, ry1 = ry0 - 1.0;
    int i = (int) p[bx0], j = (int) p[bx1];
    int b00 = (int) p[i + by0], b10 = (int) p[j + by0], b01 = (int) p[i + by1], b11 = (int) p[j + by1];
    auto sc = [] (double v) { return v * v * (3.0 - 2.0 * v); };
    auto lerp = [] (double tt, double a, double b) { return a + tt * (b - a); };
    double sx = sc (rx0), sy = sc (ry0);
    double u = rx0 * g2[b00][0] + ry0 * g2[b00][1], v = rx1 * g2[b10][0] + ry0 * g2[b10][1];
    double a = lerp (sx, u, v);
    u = rx0 * g2[b01][0] + ry1 * g2[b01][1]; v = rx1 * g2[b11][0] + ry1 * g2[b11][1];
    double b = lerp (sx, u, v);
    return 1.5 * lerp (sy, a, b);
  }
};

struct MapState {
  int width = 0, height = 0;
  std::map<std::string, double> prop;
  // GstCircleGeometricTransform precalc (gstcirclegeometrictransform.c:144-157)
  double x_center = 0.5, y_center = 0.5, radius = 0.35;
  double pcx = 0, pcy = 0, pr = 0, pr2 = 0;
  std::vector<double> sin_table, cos_table;
  Noise *noise = nullptr;
  double v[12] = { 0 };                    // element properties, in the order of ElementDef::defaults (hot loop: no lookups)
  double get (const char *n) const { return prop.at (n); }
  void circle_precalc () {
    x_center = get ("x-center"); y_center = get ("y-center"); radius = get ("radius");
    pcx = x_center * width;
    pcy = y_center * height;
    pr = radius * 0.5 * sqrt ((double) (width * width + height * height));
    pr2 = pr * pr;
  }
};

typedef void (*MapFn) (const MapState &, int, int, double *, double *);

struct ElementDef {
  const char *name;
  bool circle;
  MapFn fn;
  std::vector<std::pair<const char *, double>> defaults;   // element-specific properties
};

// gstfisheye.c:77-125
void fisheye_map (const MapState &s, int x, int y, double *ix, double *iy) {
  double width = s.width, height = s.height;
  double nx = 2.0 * x / width - 1.0, ny = 2.0 * y / height - 1.0;
  double r = sqrt ((nx * nx + ny * ny) / 2.0);
  nx *= (0.33 + 0.1 * r * r + 0.57 * pow (r, 6.0));
  ny *= (0.33 + 0.1 * r * r + 0.57 * pow (r, 6.0));
  *ix = 0.5 * (nx + 1.0) * width;
  *iy = 0.5 * (ny + 1.0) * height;
}
// gstbulge.c:132-175
void bulge_map (const MapState &s, int x, int y, double *ix, double *iy) {
  double width = s.width, height = s.height, zoom = s.v[0] /* zoom */;
  double nx = 2.0 * (x / width - s.x_center), ny = 2.0 * (y / height - s.y_center);
  double r = sqrt (0.5 * (nx * nx + ny * ny));
  double scale = 1.0 / (zoom + ((1.0 - zoom) * smoothstep (0, s.radius, r)));
  nx *= scale; ny *= scale;
  *ix = (0.5 * nx + s.x_center) * width;
  *iy = (0.5 * ny + s.y_center) * height;
}
// gstcircle.c:164-189
void circle_map (const MapState &s, int x, int y, double *ix, double *iy) {
  double dx = x - s.pcx, dy = y - s.pcy;
  double distance = sqrt (dx * dx + dy * dy);
  double theta = atan2 (-dy, -dx) + s.v[0] /* angle */;
  theta = mod_float (theta, 2 * kPi);
  *ix = s.width * theta / (s.v[1] /* spread-angle */ + 0.0001);
  *iy = s.height * (1 - (distance - s.pr) / ((int) s.v[2] /* height */ + 0.0001));
}
// gstkaleidoscope.c:165-195
void kaleidoscope_map (const MapState &s, int x, int y, double *ix, double *iy) {
  double angle = s.v[0] /* angle */, angle2 = s.v[1] /* angle2 */;
  int sides = (int) s.v[2] /* sides */;
  double dx = x - s.pcx, dy = y - s.pcy;
  double distance = sqrt (dx * dx + dy * dy);
  double theta = atan2 (dy, dx) - angle - angle2;
  theta = triangle (theta / kPi * sides * 0.5);
  if (s.pr != 0) {
    double radiusc = s.pr / cos (theta);
    distance = radiusc * triangle (distance / radiusc);
  }
  theta += angle;
  *ix = s.pcx + distance * cos (theta);
  *iy = s.pcy + distance * sin (theta);
}
// gstpinch.c:136-174
void pinch_map (const MapState &s, int x, int y, double *ix, double *iy) {
  double dx = x - s.pcx, dy = y - s.pcy;
  double distance = dx * dx + dy * dy;
  if (distance > s.pr2 || distance == 0) { *ix = x; *iy = y; return; }
  double d = sqrt (distance / s.pr2);
  double t = pow (sin (kPi * 0.5 * d), -s.v[0] /* intensity */);
  dx *= t; dy *= t;
  *ix = s.pcx + dx;
  *iy = s.pcy + dy;
}
// gstrotate.c:137-182
void rotate_map (const MapState &s, int x, int y, double *ix, double *iy) {
  int w = s.width, h = s.height;
  double ar = s.v[0] /* angle */;
  double cox = 0.5 * w, coy = 0.5 * h, cix = cox, ciy = coy;
  double xo = x - cox, yo = y - coy;
  double ao = atan2 (yo, xo);
  double r = sqrt (xo * xo + yo * yo);
  double ai = ao + ar;
  double xi = r * cos (ai), yi = r * sin (ai);
  *ix = xi + cix;
  *iy = yi + ciy;
}
// gstsphere.c:137-186
void sphere_map (const MapState &s, int x, int y, double *ix, double *iy) {
  double dx = x - s.pcx, dy = y - s.pcy;
  double dx2 = dx * dx, dy2 = dy * dy;
  if (dy2 >= (s.pr2 - (s.pr2 * dx2) / s.pr2)) { *ix = x; *iy = y; return; }
  double r_refraction = 1.0 / s.v[0] /* refraction */;
  double z = sqrt ((1.0 - dx2 / s.pr2 - dy2 / s.pr2) * (s.pr2));
  double z2 = z * z;
  double angle = acos (dx / sqrt (dx2 + z2));
  double angle1 = kPi / 2 - angle;
  double angle2 = asin (sin (angle1) * r_refraction);
  angle2 = kPi / 2 - angle - angle2;
  *ix = x - tan (angle2) * z;
  angle = acos (dy / sqrt (dy2 + z2));
  angle1 = kPi / 2 - angle;
  angle2 = asin (sin (angle1) * r_refraction);
  angle2 = kPi / 2 - angle - angle2;
  *iy = y - tan (angle2) * z;
}
// gsttwirl.c:136-164
void twirl_map (const MapState &s, int x, int y, double *ix, double *iy) {
  double dx = x - s.pcx, dy = y - s.pcy;
  double distance = dx * dx + dy * dy;
  if (distance > s.pr2) { *ix = x; *iy = y; return; }
  double d = sqrt (distance);
  double a = atan2 (dy, dx) + s.v[0] /* angle */ * (s.pr - d) / s.pr;
  *ix = s.pcx + d * cos (a);
  *iy = s.pcy + d * sin (a);
}
// gstwaterripple.c:162-195
void waterripple_map (const MapState &s, int x, int y, double *ix, double *iy) {
  double dx = x - s.pcx, dy = y - s.pcy;
  double distance = dx * dx + dy * dy;
  if (distance > s.pr2) { *ix = x; *iy = y; return; }
  double wavelength = s.v[2] /* wavelength */;
  double d = sqrt (distance);
  double amount = s.v[0] /* amplitude */ * sin (d / wavelength * kPi * 2 - s.v[1] /* phase */);
  amount *= (s.pr - d) / s.pr;
  if (d != 0) amount *= wavelength / d;
  *ix = x + dx * amount;
  *iy = y + dy * amount;
}
// gststretch.c:132-176 (MAX_SHRINK_AMOUNT 3.0, :78)
void stretch_map (const MapState &s, int x, int y, double *ix, double *iy) {
  double width = s.width, height = s.height;
  double nx = 2.0 * (x / width - s.x_center), ny = 2.0 * (y / height - s.y_center);
  double r = sqrt (0.5 * (nx * nx + ny * ny));
  double a = 1.0 + (3.0 - 1.0) * s.v[0] /* intensity */;
  double b = a - 1.0;
  nx *= a - b * smoothstep (0.0, s.radius, r);
  ny *= a - b * smoothstep (0.0, s.radius, r);
  *ix = (0.5 * nx + s.x_center) * width;
  *iy = (0.5 * ny + s.y_center) * height;
}
// gsttunnel.c:79-115 (the centre pixel divides 0/0 -> NaN -> stays unmapped)
void tunnel_map (const MapState &s, int x, int y, double *ix, double *iy) {
  double width = s.width, height = s.height;
  double m = width > height ? width : height;
  double nx = 2.0 * (x - s.x_center * width) / m, ny = 2.0 * (y - s.y_center * height) / m;
  double r = sqrt (0.5 * (nx * nx + ny * ny));
  nx *= clampd (r, 0.0, s.radius) / r;
  ny *= clampd (r, 0.0, s.radius) / r;
  *ix = 0.5 * (nx) * m + s.x_center * width;
  *iy = 0.5 * (ny) * m + s.y_center * height;
}
// gstsquare.c:158-192
void square_map (const MapState &s, int x, int y, double *ix, double *iy) {
  double width = s.width, height = s.height;
  double sw = s.v[0] /* width */, sh = s.v[1] /* height */, zoom = s.v[2] /* zoom */;
  double nx = 2.0 * x / width - 1.0, ny = 2.0 * y / height - 1.0;
  double ax = nx < 0 ? -nx : nx, ay = ny < 0 ? -ny : ny;
  nx *= (1.0 / zoom) * (1.0 + (zoom - 1.0) * smoothstep (sw - 0.125, sw + 0.125, ax));
  ny *= (1.0 / zoom) * (1.0 + (zoom - 1.0) * smoothstep (sh - 0.125, sh + 0.125, ay));
  *ix = 0.5 * (nx + 1.0) * width;
  *iy = 0.5 * (ny + 1.0) * height;
}
// gstmirror.c:158-203 (mode: 0 left, 1 right, 2 top, 3 bottom; gstmirror.h:77-83)
void mirror_map (const MapState &s, int x, int y, double *ix, double *iy) {
  double hw = s.width / 2.0 - 1.0, hh = s.height / 2.0 - 1.0;
  switch ((int) s.v[0] /* mode */) {
    case 0: *ix = (x > hw) ? s.width - 1.0 - x : x; *iy = y; break;
    case 1: *ix = (x > hw) ? x : s.width - 1.0 - x; *iy = y; break;
    case 2: *iy = (y > hh) ? s.height - 1.0 - y : y; *ix = x; break;
    default: *iy = (y > hh) ? y : s.height - 1.0 - y; *ix = x; break;
  }
}
// gstperspective.c:184-210
void perspective_map (const MapState &s, int x, int y, double *ix, double *iy) {
  double m[9];
  for (int i = 0; i < 9; i++) m[i] = s.v[i];     // matrix-0 .. matrix-8, row-major
  double xp = (m[0] * x + m[1] * y + m[2]);
  double yp = (m[3] * x + m[4] * y + m[5]);
  double w = (m[6] * x + m[7] * y + m[8]);
  *ix = xp / w;
  *iy = yp / w;
}
// gstmarble.c:185-227 (noise unpinned, see Noise above)
void marble_map (const MapState &s, int x, int y, double *ix, double *iy) {
  double xscale = s.v[0] /* x-scale */;
  int displacement = (int) (127 * (1 + s.noise->at (x / xscale, y / xscale)));
  displacement = displacement > 255 ? 255 : (displacement < 0 ? 0 : displacement);
  *ix = x + s.sin_table[displacement];
  *iy = y + s.cos_table[displacement];
}

const std::vector<ElementDef> &elements () {
  static const std::vector<ElementDef> defs = {
    { "fisheye", false, fisheye_map, {} },
    { "bulge", true, bulge_map, { { "zoom", 3.0 } } },
    { "circle", true, circle_map, { { "angle", 0.0 }, { "spread-angle", kPi }, { "height", 20.0 } } },
    { "kaleidoscope", true, kaleidoscope_map, { { "angle", 0.0 }, { "angle2", 0.0 }, { "sides", 3.0 } } },
    { "pinch", true, pinch_map, { { "intensity", 0.5 } } },
    { "rotate", false, rotate_map, { { "angle", 0.0 } } },
    { "sphere", true, sphere_map, { { "refraction", 1.5 } } },
    { "twirl", true, twirl_map, { { "angle", kPi } } },
    { "waterripple", true, waterripple_map, { { "amplitude", 10.0 }, { "phase", 0.0 }, { "wavelength", 16.0 } } },
    { "stretch", true, stretch_map, { { "intensity", 0.5 } } },
    { "tunnel", true, tunnel_map, {} },
    { "square", false, square_map, { { "width", 0.5 }, { "height", 0.5 }, { "zoom", 2.0 } } },
    { "mirror", false, mirror_map, { { "mode", 0.0 } } },
    { "perspective", false, perspective_map, { { "matrix-0", 1 }, { "matrix-1", 0 }, { "matrix-2", 0 }, { "matrix-3", 0 },
        { "matrix-4", 1 }, { "matrix-5", 0 }, { "matrix-6", 0 }, { "matrix-7", 0 }, { "matrix-8", 1 } } },
    { "marble", false, marble_map, { { "x-scale", 4.0 }, { "y-scale", 4.0 }, { "amount", 1.0 }, { "turbulence", 1.0 } } },
  };
  return defs;
}

// element + properties -> MapState (properties validated like a GObject setter would)
struct BuiltState {
  MapState s;
  const ElementDef *def = nullptr;
  Noise *noise = nullptr;
  ~BuiltState () { delete noise; }
};
int make_state (BuiltState &b, const char *who, const char *element, int width, int height, const char *const *prop_names,
    const double *prop_values, int nprops)
{
  for (const auto &d : elements ()) if (!strcmp (d.name, element)) b.def = &d;
  if (!b.def) {
    if (!strcmp (element, "diffuse"))
      b200vf_set_error ("%s: `diffuse` draws a fresh random map every frame (gstdiffuse.c:151-189); "
          "it has no precalculated map and no bit-exact counterpart", who);
    else
      b200vf_set_error ("%s: unknown element `%s`", who, element);
    return B200VF_E_UNSUPPORTED;
  }
  MapState &s = b.s;
  s.width = width; s.height = height;
  for (const auto &kv : b.def->defaults) s.prop[kv.first] = kv.second;
  if (b.def->circle) { s.prop["x-center"] = 0.5; s.prop["y-center"] = 0.5; s.prop["radius"] = 0.35; }
  for (int i = 0; i < nprops; i++) {
    B200VF_REQUIRE (prop_names && prop_values && prop_names[i], B200VF_E_INVAL, "gt map: NULL property");
    if (!strcmp (prop_names[i], "off-edge-pixels")) continue;    // a gather policy, not a map input
    auto it = s.prop.find (prop_names[i]);
    if (it == s.prop.end ()) {
      b200vf_set_error ("%s: element `%s` has no property `%s`", who, element, prop_names[i]);
      return B200VF_E_PROPERTY;
    }
    it->second = prop_values[i];
  }
  if (b.def->circle) s.circle_precalc ();
  for (size_t i = 0; i < b.def->defaults.size (); i++) s.v[i] = s.prop[b.def->defaults[i].first];
  if (!strcmp (element, "marble")) {                             // marble_prepare, gstmarble.c:160-183
    b.noise = new Noise (0x9e3779b9u);
    s.noise = b.noise;
    s.sin_table.resize (256); s.cos_table.resize (256);
    for (int i = 0; i < 256; i++) {
      double angle = (kPi * 2 * i) / 256.0 * s.get ("turbulence");
      s.sin_table[i] = -s.get ("y-scale") * sin (angle);
      s.cos_table[i] = s.get ("y-scale") * cos (angle);
    }
  }
  return B200VF_OK;
}

// do_map's policy, truncation and bounds test for one pixel (gstgeometrictransform.c:167-207)
inline int32_t resolve_one (double in_x, double in_y, int width, int height, int off_edge) {
  if (off_edge == 1) {
    in_x = clampd (in_x, 0, width - 1);
    in_y = clampd (in_y, 0, height - 1);
  } else if (off_edge == 2) {
    in_x = mod_float (in_x, width);
    in_y = mod_float (in_y, height);
    if (in_x < 0) in_x += width;
    if (in_y < 0) in_y += height;
  }
  int tx = (int) in_x, ty = (int) in_y;    // NaN / out-of-range -> INT_MIN on x86-64, like the reference build
  return (tx >= 0 && tx < width && ty >= 0 && ty < height) ? ty * width + tx : -1;
}

}  // namespace

B200VF_API int b200vf_gt_build_map (const char *element, int width, int height, const char *const *prop_names,
    const double *prop_values, int nprops, double *map_xy)
{
  B200VF_REQUIRE (element && map_xy && width > 0 && height > 0 && nprops >= 0, B200VF_E_INVAL, "gt_build_map: bad argument");
  BuiltState b;
  int rc = make_state (b, "gt_build_map", element, width, height, prop_names, prop_values, nprops);
  if (rc) return rc;
  const MapState &s = b.s;
  const ElementDef *def = b.def;
  // every pixel is independent: rows are split over the host cores (the reference builds its map
  // on one thread under the object lock, 3.6 s at 8K; the values do not depend on the split)
  unsigned nthreads = std::thread::hardware_concurrency ();
  if (nthreads > 32) nthreads = 32;
  if (nthreads < 1 || (long long) width * height < (1 << 18)) nthreads = 1;
  auto rows = [&] (int y0, int y1) {
    for (int y = y0; y < y1; y++) {
      double *ptr = map_xy + (size_t) y * width * 2;
      for (int x = 0; x < width; x++, ptr += 2) def->fn (s, x, y, ptr, ptr + 1);
    }
  };
  if (nthreads == 1) rows (0, height);
  else {
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < nthreads; t++)
      pool.emplace_back (rows, (int) ((long long) height * t / nthreads), (int) ((long long) height * (t + 1) / nthreads));
    for (auto &th : pool) th.join ();
  }
  return B200VF_OK;
}

// Internal (gt_device_maps.cu): the host's answer for a list of pixels - the entries the GPU evaluation could not
// certify (a coordinate within its libm error bound of an integer).
int b200vf_gt_host_index_at (const char *element, int width, int height, const char *const *prop_names, const double *prop_values,
    int nprops, int off_edge, const int32_t *pixels, size_t n, int32_t *index_out)
{
  BuiltState b;
  int rc = make_state (b, "gt_host_index_at", element, width, height, prop_names, prop_values, nprops);
  if (rc) return rc;
  for (size_t i = 0; i < n; i++) {
    const int x = pixels[i] % width, y = pixels[i] / width;
    double ix, iy;
    b.def->fn (b.s, x, y, &ix, &iy);
    index_out[i] = resolve_one (ix, iy, width, height, off_edge);
  }
  return B200VF_OK;
}

// Internal (gt_device_maps.cu): marble's host-built tables (noise lattice from the fixed seed, sin / cos displacement
// tables from glibc): p[514], g2[514][2], sin[256], cos[256] = 2054 doubles. The map over them needs no libm.
int b200vf_gt_marble_tables (const char *const *prop_names, const double *prop_values, int nprops, double *out)
{
  BuiltState b;
  int rc = make_state (b, "gt_marble_tables", "marble", 8, 8, prop_names, prop_values, nprops);
  if (rc) return rc;
  memcpy (out, b.noise->p, sizeof b.noise->p);
  memcpy (out + 514, b.noise->g2, sizeof b.noise->g2);
  memcpy (out + 514 + 1028, b.s.sin_table.data (), 256 * sizeof (double));
  memcpy (out + 514 + 1028 + 256, b.s.cos_table.data (), 256 * sizeof (double));
  return B200VF_OK;
}

// do_map (gstgeometrictransform.c:167-207) minus the memcpy: off-edge policy, truncation
// toward zero, bounds test. index = ty*width+tx, or -1 when the output keeps the fill.
B200VF_API int b200vf_gt_resolve_map (const double *map_xy, int width, int height, int off_edge, int32_t *index_out)
{
  B200VF_REQUIRE (map_xy && index_out && width > 0 && height > 0, B200VF_E_INVAL, "gt_resolve_map: bad argument");
  B200VF_REQUIRE (off_edge >= 0 && off_edge <= 2, B200VF_E_PROPERTY, "gt_resolve_map: off-edge-pixels %d", off_edge);
  B200VF_REQUIRE ((long long) width * height < 0x7fffffffll, B200VF_E_INVAL, "gt_resolve_map: frame too large for int32 indices");
  const double *ptr = map_xy;
  for (long long i = 0, n = (long long) width * height; i < n; i++, ptr += 2)
    index_out[i] = resolve_one (ptr[0], ptr[1], width, height, off_edge);
  return B200VF_OK;
}

B200VF_API int b200vf_gt_index_row_range (const int32_t *index, size_t n, int width, int *row_lo, int *row_hi) {
  B200VF_REQUIRE (index && row_lo && row_hi && width > 0, B200VF_E_INVAL, "gt_index_row_range: bad argument");
  int32_t lo = INT32_MAX, hi = -1;
  for (size_t i = 0; i < n; i++) {
    int32_t v = index[i];
    if (v < 0) continue;
    if (v < lo) lo = v;
    if (v > hi) hi = v;
  }
  if (hi < 0) { *row_lo = *row_hi = 0; return B200VF_OK; }
  *row_lo = lo / width;
  *row_hi = hi / width + 1;
  return B200VF_OK;
}
