/* oracle_port.c - CPU restatement ("port") of the gst-plugins-bad 1.19.2
 * per-pixel video-filter hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the checker: tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * are the only callers.  The product (gst-plugins-bad_b200/) never links,
 * imports or executes it; there is no CPU fallback in the product.
 *
 * Parity status: PINNED.  The reference's tests hold no golden vectors for
 * these elements, so every function here is checked bit-exactly
 * against the reference's own C compiled from /root/reference by
 * oracle/build_ref.py (tests/test_oracle_vs_ref.py, run wherever oracle/_ref
 * exists) and against the golden frames that library generated
 * (tests/golden/, made by tests/golden/make_golden.py).
 *
 * Style: closed forms rather than the reference's ring-buffer / ORC program
 * structure; each function cites the reference lines it restates
 * (paths relative to /root/reference).  Compile with -O2 -ffp-contract=off.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <limits.h>

#define EXPORT __attribute__((visibility("default")))

/* ------------------------------------------------------------------ bayer2rgb
 * gst/bayer/gstbayer2rgb.c:354-451 + the ORC programs
 * gst/bayer/gstbayerorc.orc:3-248 (C backups gstbayerorc-dist.c).            */

static inline uint8_t avgub (uint8_t a, uint8_t b) { return (uint8_t) ((a + b + 1) >> 1); }

/* One Bayer row -> the two horizontally upsampled lines (gstbayer2rgb.c:354-381):
 * h0 carries the even-column samples, h1 the odd-column samples. */
static void
bayer_upsample_row (const uint8_t *s, int n, uint8_t *h0, uint8_t *h1)
{
  for (int x = 0; x < n; x++) {
    if (x & 1) {
      h1[x] = s[x];
      h0[x] = (x + 1 < n) ? avgub (s[x - 1], s[x + 1]) : s[x - 1];
    } else {
      h0[x] = s[x];
      h1[x] = (x == 0) ? s[1] : avgub (s[x - 1], s[x + 1]);
    }
  }
  /* right-edge rule (:372-380): the last two columns copy, they never average */
  h1[n - 2] = s[n - 3];
  h0[n - 1] = s[n - 2];
}

/* format: 0 bggr, 1 gbrg, 2 grbg, 3 rggb (enum gstbayer2rgb.c:95-101).
 * Domain: even width >= 4, height >= 3 (the reference reads stale or
 * out-of-row memory outside it). */
EXPORT int
oracle_bayer2rgb (uint8_t *dest, int dest_stride, const uint8_t *src, int src_stride,
    int width, int height, int format, int r_off, int g_off, int b_off)
{
  if (width < 4 || (width & 1) || height < 3)
    return -1;
  /* functions are named for BGGR; RGGB and GBRG swap red/blue (:399-407),
   * GRBG and GBRG start with the "gr" row (:422-427) */
  if (format == 3 || format == 1) { int t = r_off; r_off = b_off; b_off = t; }
  int first_is_gr = (format == 2 || format == 1);
  int a_off = 6 - r_off - g_off - b_off;        /* the remaining byte gets 255 */

  uint8_t *lines = malloc ((size_t) 6 * width);
  uint8_t *h0u = lines, *h1u = h0u + width, *h0c = h1u + width, *h1c = h0c + width,
      *h0d = h1c + width, *h1d = h0d + width;

  for (int j = 0; j < height; j++) {
    int ju = (j == 0) ? 1 : j - 1;              /* top edge mirrors row 1 (:432-433) */
    /* bottom edge: the ring slot of row `height` still holds row height-4
     * (or the preloaded row 1 when height == 3) (:429-448) */
    int jd = (j + 1 < height) ? j + 1 : (height >= 4 ? height - 4 : 1);
    bayer_upsample_row (src + (size_t) ju * src_stride, width, h0u, h1u);
    bayer_upsample_row (src + (size_t) j * src_stride, width, h0c, h1c);
    bayer_upsample_row (src + (size_t) jd * src_stride, width, h0d, h1d);
    uint8_t *d = dest + (size_t) j * dest_stride;
    int gr_row = ((j & 1) != 0) ^ first_is_gr;
    for (int x = 0; x < width; x++) {
      uint8_t r, g, b;
      if (!gr_row) {                            /* bayer_orc_merge_bg_* (.orc:43-66) */
        b = h0c[x];
        r = avgub (h1u[x], h1d[x]);
        g = (x & 1) ? h1c[x] : avgub (avgub (h0u[x], h0d[x]), h1c[x]);
      } else {                                  /* bayer_orc_merge_gr_* (.orc:69-92) */
        r = h1c[x];
        b = avgub (h0u[x], h0d[x]);
        g = (x & 1) ? avgub (avgub (h1u[x], h1d[x]), h0c[x]) : h0c[x];
      }
      d[4 * x + r_off] = r;
      d[4 * x + g_off] = g;
      d[4 * x + b_off] = b;
      d[4 * x + a_off] = 255;
    }
  }
  free (lines);
  return 0;
}

/* rgb2bayer: gst/bayer/gstrgb2bayer.c:254-267 — restated with
 * the same argument convention as the product entry point. src is ARGB. */
EXPORT int
oracle_rgb2bayer (uint8_t *dest, int dest_stride, const uint8_t *src, int src_stride,
    int width, int height, int format)
{
  /* format: 0 bggr, 1 gbrg, 2 grbg, 3 rggb (gstrgb2bayer.h) */
  for (int j = 0; j < height; j++) {
    for (int i = 0; i < width; i++) {
      int is_blue = ((j & 1) << 1) | (i & 1);
      const uint8_t *p = src + (size_t) j * src_stride + 4 * i;
      uint8_t v;
      if (is_blue == format) v = p[3];
      else if ((is_blue ^ 3) == format) v = p[1];
      else v = p[2];
      dest[(size_t) j * dest_stride + i] = v;
    }
  }
  return 0;
}

/* --------------------------------------------------------------- gaussianblur
 * gst/gaudieffects/gstgaussblur.c:259-422                                   */

/* make_gaussian_kernel (:361-422). Returns windowsize; kernel/kernel_sum must
 * hold 2*ceil(2.5*|sigma|)+1 floats (<= 101 for |sigma| <= 20). */
EXPORT int
oracle_gauss_kernel (float sigma, float *kernel, float *kernel_sum)
{
  const float fe = -0.5 / (sigma * sigma);
  const float dx = 1.0 / (sigma * sqrt (2 * 3.1415926535897932384626433832795028841971693993751));
  int center = (int) ceil (2.5 * fabs (sigma));
  int ws = 1 + 2 * center;
  if (ws == 1) { kernel[0] = 1.0f; kernel_sum[0] = 1.0f; return 1; }
  float sum = kernel[center] = dx;
  for (int i = 1; i <= center; i++) {
    float fx = dx * pow (2.7182818284590452353602874713526624977572470937000, fe * i * i);
    kernel[center + i] = kernel[center - i] = fx;
    sum += 2 * fx;
  }
  if (sigma < 0) {              /* negative sigma sharpens (:395-398) */
    sum = -sum;
    kernel[center] += 2.0 * sum;
  }
  for (int i = 0; i < ws; i++) kernel[i] /= sum;
  float acc = 0.0f;
  for (int i = 0; i < ws; i++) { acc += kernel[i]; kernel_sum[i] = acc; }
  return ws;
}

/* Window of taps for output index `pos` in a line of `len` samples
 * (blur_row_x :266-275 and gaussian_smooth :313-322): taps k in [kmin,kmax)
 * read samples first+0 .., and the normaliser is the partial kernel sum. */
static inline void
gauss_window (int pos, int len, int ws, const float *ksum, int *kmin, int *kmax, int *first, float *sum)
{
  int center = ws / 2;
  int cc = center - pos;
  *kmin = cc > 0 ? cc : 0;
  *first = *kmin - cc;
  *kmax = ws < len - *first ? ws : len - *first;
  float s = ksum[*kmax - 1];
  s -= *kmin ? ksum[*kmin - 1] : 0.0;
  *sum = s;
}

/* image/out point at COMP_DATA(frame,0) = plane + p0; the caller
 * has already copied in -> out (gst_video_frame_copy, :252). */
EXPORT int
oracle_gaussblur (const uint8_t *image, uint8_t *out, int width, int height, int stride,
    const float *kernel, const float *kernel_sum, int ws)
{
  float *tmp = malloc (sizeof (float) * (size_t) stride * height + 64);
  if (!tmp) return -1;
  /* horizontal pass (blur_row_x) */
  for (int r = 0; r < height; r++) {
    const uint8_t *in_row = image + (size_t) r * stride;
    float *out_row = tmp + (size_t) r * stride;
    for (int c = 0; c < width; c++) {
      int kmin, kmax, first; float sum;
      gauss_window (c, width, ws, kernel_sum, &kmin, &kmax, &first, &sum);
      float dot[4] = { 0.0f, 0.0f, 0.0f, 0.0f };
      const uint8_t *p = in_row + 4 * first;
      for (int k = kmin; k < kmax; k++, p += 4) {
        float coeff = kernel[k];
        for (int ch = 0; ch < 4; ch++) dot[ch] += (float) p[ch] * coeff;
      }
      for (int ch = 0; ch < 4; ch++) out_row[4 * c + ch] = dot[ch] / sum;
    }
  }
  /* vertical pass (gaussian_smooth :336-354) */
  for (int r = 0; r < height; r++) {
    int kmin, kmax, first; float sum;
    gauss_window (r, height, ws, kernel_sum, &kmin, &kmax, &first, &sum);
    uint8_t *out_row = out + (size_t) r * stride;
    for (int c = 0; c < width; c++) {
      float dot[4] = { 0.0f, 0.0f, 0.0f, 0.0f };
      const float *t = tmp + (size_t) first * stride + 4 * c;
      for (int k = kmin; k < kmax; k++, t += stride) {
        float kern = kernel[k];
        for (int ch = 0; ch < 4; ch++) dot[ch] += t[ch] * kern;
      }
      for (int ch = 0; ch < 4; ch++) {
        double v = dot[ch] / sum + 0.5;         /* fp32 divide, then double +0.5 (:348) */
        v = v > 255 ? 255 : (v < 0 ? 0 : v);
        out_row[4 * c + ch] = (uint8_t) v;
      }
    }
  }
  free (tmp);
  return 0;
}

/* ------------------------------------------------------ gaudieffects point ops */

/* burn: gaudi_orc_burn, gst/gaudieffects/gstgaudieffectsorc.orc:1-25
 * (C backup gstgaudieffectsorc-dist.c:147-265). All four bytes transformed. */
static inline uint8_t
burn_byte (uint8_t c, int adj)
{
  uint16_t a = (uint16_t) ((uint16_t) (c + adj) >> 1);
  uint16_t t = (uint16_t) ((uint8_t) (255 - c)) << 7;
  unsigned q;
  if ((a & 0xff) == 0) q = 255;
  else { q = t / (a & 0xff); if (q > 255) q = 255; }
  return (uint8_t) (255 - q);
}

EXPORT void
oracle_burn (uint32_t *dest, const uint32_t *src, int adjustment, int n)
{
  for (int i = 0; i < n; i++) {
    uint32_t in = src[i], o = 0;
    for (int b = 0; b < 4; b++)
      o |= (uint32_t) burn_byte ((in >> (8 * b)) & 0xff, adjustment) << (8 * b);
    dest[i] = o;
  }
}

static inline int clamp255 (int v) { return v > 255 ? 255 : (v < 0 ? 0 : v); }

/* dodge: gst/gaudieffects/gstdodge.c:231-254 */
EXPORT void
oracle_dodge (const uint32_t *src, uint32_t *dest, int n)
{
  for (int i = 0; i < n; i++) {
    uint32_t in = src[i], o = 0;
    for (int sh = 0; sh <= 16; sh += 8) {
      int v = (in >> sh) & 0xff;
      o |= (uint32_t) clamp255 ((256 * v) / (256 - v)) << sh;
    }
    dest[i] = o;                                /* byte 3 := 0 (:252) */
  }
}

/* chromium: gst/gaudieffects/gstchromium.c:102-110,282-338 */
EXPORT void
oracle_chromium (const uint32_t *src, uint32_t *dest, int n, int edge_a, int edge_b)
{
  static int table[1024];
  const float pi = 3.141582f;                   /* sic (:102) */
  for (int angle = 0; angle < 1024; angle++) {
    float rad = ((float) angle / 512) * pi;
    table[angle] = (int) (cos (rad) * 512);
  }
  for (int i = 0; i < n; i++) {
    uint32_t in = src[i], o = 0;
    for (int sh = 0; sh <= 16; sh += 8) {
      int v = (in >> sh) & 0xff;
      int c = table[((v + edge_a) + ((v * edge_b) / 2)) & 1023];
      if (c < 0) c = -c;
      o |= (uint32_t) clamp255 (c) << sh;
    }
    dest[i] = o;
  }
}

/* exclusion: gst/gaudieffects/gstexclusion.c:256-284 (red uses green*red, :269-270) */
EXPORT void
oracle_exclusion (const uint32_t *src, uint32_t *dest, int n, int factor)
{
  for (int i = 0; i < n; i++) {
    uint32_t in = src[i];
    int red = (in >> 16) & 0xff, green = (in >> 8) & 0xff, blue = in & 0xff;
    int r2 = factor - (((factor - red) * (factor - red) / factor) + ((green * red) / factor));
    int g2 = factor - (((factor - green) * (factor - green) / factor) + ((green * green) / factor));
    int b2 = factor - (((factor - blue) * (factor - blue) / factor) + ((blue * blue) / factor));
    dest[i] = ((uint32_t) clamp255 (r2) << 16) | ((uint32_t) clamp255 (g2) << 8) | (uint32_t) clamp255 (b2);
  }
}

/* solarize: gst/gaudieffects/gstsolarize.c:286-339. The mixed gint / guint32
 * arithmetic (:316-327) is kept type-for-type. */
static inline uint32_t
solarize_channel (uint32_t v, int start, int period, int up_length, int down_length)
{
  static const unsigned int ceiling = 255;
  uint32_t color;
  int param = (int) v;
  param += 256;
  param -= start;
  param %= period;
  if (param < up_length) {
    color = param * ceiling;
    color /= up_length;
  } else {
    color = down_length - (param - up_length);
    color *= ceiling;
    color /= down_length;
  }
  return color > 255 ? 255 : color;
}

EXPORT void
oracle_solarize (const uint32_t *src, uint32_t *dest, int n, int threshold, int start, int end)
{
  int period = 1, up_length = 1, down_length = 1;
  if (end != start) period = end - start;
  if (threshold != start) up_length = threshold - start;
  if (threshold != end) down_length = end - threshold;
  for (int i = 0; i < n; i++) {
    uint32_t in = src[i], o = 0;
    for (int sh = 0; sh <= 16; sh += 8)
      o |= solarize_channel ((in >> sh) & 0xff, start, period, up_length, down_length) << sh;
    dest[i] = o;
  }
}

/* dilate: gst/gaudieffects/gstdilate.c:258-345. Candidates in the order
 * down, right, up, left with strict compare; `up` is always the pixel itself
 * (:291-294, `up < src` is always true) so it can never win. */
static inline uint32_t
dilate_lum (uint32_t in)
{
  return 90 * ((in >> 16) & 0xff) + 115 * ((in >> 8) & 0xff) + 51 * (in & 0xff);
}

EXPORT void
oracle_dilate (const uint32_t *src, uint32_t *dest, int width, int height, int erode)
{
  for (int y = 0; y < height; y++) {
    for (int x = 0; x < width; x++) {
      const uint32_t *p = src + (size_t) y * width + x;
      uint32_t best = *p, bl = dilate_lum (best);
      uint32_t cand[3];
      cand[0] = (y + 1 < height) ? p[width] : *p;       /* down */
      cand[1] = (x + 1 < width) ? p[1] : *p;            /* right */
      cand[2] = (x > 0) ? p[-1] : *p;                   /* left (after the dead `up`) */
      for (int k = 0; k < 3; k++) {
        uint32_t l = dilate_lum (cand[k]);
        if (erode ? (l < bl) : (l > bl)) { best = cand[k]; bl = l; }
      }
      dest[(size_t) y * width + x] = best;
    }
  }
}

/* --------------------------------------------------------------- coloreffects
 * gst/coloreffects/gstcoloreffects.c:288-435. `table` = one of the five
 * 256x3 preset tables (:117-286), map_luma per preset (:503-548). In place. */
static const int ycbcr_to_rgb[12] = { 298, 0, 409, -57068, 298, -100, -208, 34707, 298, 516, 0, -70870 };
static const int rgb_to_ycbcr[12] = { 66, 129, 25, 4096, -38, -74, 112, 32768, 112, -94, -18, 32768 };
static inline int
mat (const int *m, int o, int a, int b, int c)
{
  return (m[o * 4] * a + m[o * 4 + 1] * b + m[o * 4 + 2] * c + m[o * 4 + 3]) >> 8;
}

EXPORT void
oracle_coloreffects (uint8_t *data, int width, int height, int stride, int pstride,
    int o0, int o1, int o2, const uint8_t *table, int map_luma, int is_ayuv)
{
  if (!table) return;                           /* preset none (:488-490) */
  for (int i = 0; i < height; i++) {
    uint8_t *p = data + (size_t) i * stride;
    for (int j = 0; j < width; j++, p += pstride) {
      if (!is_ayuv) {                           /* transform_rgb :303-359 */
        uint32_t r = p[o0], g = p[o1], b = p[o2];
        if (map_luma) {
          uint32_t luma = ((r << 8) * 54) + ((g << 8) * 183) + ((b << 8) * 19);
          luma >>= 16;
          luma *= 3;
          p[o0] = table[luma]; p[o1] = table[luma + 1]; p[o2] = table[luma + 2];
        } else {
          p[o0] = table[r * 3]; p[o1] = table[g * 3 + 1]; p[o2] = table[b * 3 + 2];
        }
      } else {                                  /* transform_ayuv :361-435 */
        int y = p[o0], u = p[o1], v = p[o2], r, g, b;
        if (map_luma) {
          r = table[y * 3]; g = table[y * 3 + 1]; b = table[y * 3 + 2];
        } else {
          r = clamp255 (mat (ycbcr_to_rgb, 0, y, u, v));
          g = clamp255 (mat (ycbcr_to_rgb, 1, y, u, v));
          b = clamp255 (mat (ycbcr_to_rgb, 2, y, u, v));
          r = table[r * 3]; g = table[g * 3 + 1]; b = table[b * 3 + 2];
        }
        p[o0] = (uint8_t) clamp255 (mat (rgb_to_ycbcr, 0, r, g, b));
        p[o1] = (uint8_t) clamp255 (mat (rgb_to_ycbcr, 1, r, g, b));
        p[o2] = (uint8_t) clamp255 (mat (rgb_to_ycbcr, 2, r, g, b));
      }
    }
  }
}

/* ----------------------------------------------------------------- chromahold
 * gst/coloreffects/gstchromahold.c:271-360 */
EXPORT int
oracle_rgb_to_hue (int r, int g, int b)
{
  int m = r < g ? r : g; if (b < m) m = b;
  int M = r > g ? r : g; if (b > M) M = b;
  int C = M - m, C2 = C >> 1, h;
  if (C == 0) return (int) UINT_MAX;            /* G_MAXUINT -> -1 (:282) */
  else if (M == r) h = ((256 * 60 * (g - b) + C2) / C);
  else if (M == g) h = ((256 * 60 * (b - r) + C2) / C) + 120 * 256;
  else h = ((256 * 60 * (r - g) + C2) / C) + 240 * 256;
  h >>= 8;
  if (h >= 360) h -= 360; else if (h < 0) h += 360;
  return h;
}

EXPORT void
oracle_chromahold (uint8_t *data, int width, int height, int stride,
    int pr, int pg, int pb, int target_r, int target_g, int target_b, int tolerance)
{
  int h1 = oracle_rgb_to_hue (target_r, target_g, target_b);
  for (int i = 0; i < height; i++) {
    uint8_t *p = data + (size_t) i * stride;
    for (int j = 0; j < width; j++, p += 4) {
      int r = p[pr], g = p[pg], b = p[pb];
      int h2 = oracle_rgb_to_hue (r, g, b);
      int d1 = h1 - h2, d2 = h2 - h1;
      if (d1 < 0) d1 += 360;
      if (d2 < 0) d2 += 360;
      int diff = d1 < d2 ? d1 : d2;
      if (h1 == -1 || diff > tolerance) {
        int grey = clamp255 ((13938 * r + 46869 * g + 4730 * b) >> 16);
        p[pr] = p[pg] = p[pb] = (uint8_t) grey;
      }
    }
  }
}

/* --------------------------------------------------------- geometrictransform
 * gst/geometrictransform/gstgeometrictransform.c:167-207 (do_map) and
 * :226-293 (transform_frame). The double (x,y) map is an input here; the
 * map functions are restated in oracle_port_maps.c. */
static double
mod_float (double a, double b)                  /* geometricmath.c:171-180 */
{
  int n = (int) (a / b);
  a -= n * b;
  if (a < 0) return a + b;
  return a;
}

/* off_edge: 0 ignore, 1 clamp, 2 wrap (enum gstgeometrictransform.c:57-75) */
EXPORT void
oracle_remap (const uint8_t *in, uint8_t *out, size_t out_size, const double *map,
    int width, int height, int pixel_stride, int row_stride, int off_edge, int is_ayuv)
{
  if (is_ayuv) {
    for (size_t i = 0; i + 4 <= out_size; i += 4) {
      out[i] = 0xff; out[i + 1] = 0x10; out[i + 2] = 0x80; out[i + 3] = 0x80;
    }
  } else memset (out, 0, out_size);
  for (int y = 0; y < height; y++)
    for (int x = 0; x < width; x++, map += 2) {
      double in_x = map[0], in_y = map[1];
      if (off_edge == 1) {
        in_x = in_x > width - 1 ? width - 1 : (in_x < 0 ? 0 : in_x);
        in_y = in_y > height - 1 ? height - 1 : (in_y < 0 ? 0 : in_y);
      } else if (off_edge == 2) {
        in_x = mod_float (in_x, width);
        in_y = mod_float (in_y, height);
        if (in_x < 0) in_x += width;
        if (in_y < 0) in_y += height;
      }
      int tx = (int) in_x, ty = (int) in_y;
      if (tx >= 0 && tx < width && ty >= 0 && ty < height)
        memcpy (out + (size_t) y * row_stride + (size_t) x * pixel_stride,
            in + (size_t) ty * row_stride + (size_t) tx * pixel_stride, pixel_stride);
    }
}

/* fisheye_map: gst/geometrictransform/gstfisheye.c:77-125 */
EXPORT void
oracle_map_fisheye (double *map, int width, int height)
{
  double w = width, h = height;
  for (int y = 0; y < height; y++)
    for (int x = 0; x < width; x++, map += 2) {
      double nx = 2.0 * x / w - 1.0, ny = 2.0 * y / h - 1.0;
      double r = sqrt ((nx * nx + ny * ny) / 2.0);
      nx *= (0.33 + 0.1 * r * r + 0.57 * pow (r, 6.0));
      ny *= (0.33 + 0.1 * r * r + 0.57 * pow (r, 6.0));
      map[0] = 0.5 * (nx + 1.0) * w;
      map[1] = 0.5 * (ny + 1.0) * h;
    }
}

/* --------------------------------------------------- videofilters
 * zebrastripe: gst/videofilters/gstzebrastripe.c:242-250 (loop), :150-151 (threshold).
 * luma = first luma byte (plane 0 + offset + y_position), ps = pixel stride. */
EXPORT int
oracle_zebrastripe_y_threshold (int threshold)
{
  return 16 + (int) floor (0.5 + 2.19 * threshold);
}

EXPORT void
oracle_zebrastripe (uint8_t *luma, int ps, int stride, int width, int height, int y_threshold, int t)
{
  for (int j = 0; j < height; j++)
    for (int i = 0; i < width; i++) {
      uint8_t *p = luma + (size_t) j * stride + (size_t) i * ps;
      if (*p >= y_threshold && ((i + j + t) & 0x4)) *p = 16;
    }
}

/* videodiff luma loop: gst/videofilters/gstvideodiff.c:101-117 */
EXPORT void
oracle_videodiff_luma (uint8_t *out, const uint8_t *cur, const uint8_t *old, int stride, int width, int height,
    int threshold, int t)
{
  for (int j = 0; j < height; j++)
    for (int i = 0; i < width; i++) {
      int s1 = old[(size_t) j * stride + i], s2 = cur[(size_t) j * stride + i];
      uint8_t v = (uint8_t) s2;
      if ((s2 < s1 - threshold) || (s2 > s1 + threshold)) v = ((i + j + t) & 0x4) ? 16 : 240;
      out[(size_t) j * stride + i] = v;
    }
}

/* orc_sad_nxm_u8 (`accsadubl`, gstscenechangeorc.orc; 32-bit accumulator) */
EXPORT uint32_t
oracle_sad_u8 (const uint8_t *a, int a_stride, const uint8_t *b, int b_stride, int width, int height)
{
  uint32_t acc = 0;
  for (int j = 0; j < height; j++)
    for (int i = 0; i < width; i++) {
      int d = (int) a[(size_t) j * a_stride + i] - (int) b[(size_t) j * b_stride + i];
      acc += (uint32_t) (d < 0 ? -d : d);
    }
  return acc;
}

/* scenechange decision, gst/videofilters/gstscenechange.c:196-236.
 * state = { int n_diffs; double diffs[5]; } (the element's fields, :43-44) */
typedef struct { int n_diffs; double diffs[5]; } oracle_scenechange;
EXPORT int
oracle_scenechange_update (oracle_scenechange *sc, double score)
{
  int change;
  memmove (sc->diffs, sc->diffs + 1, sizeof (double) * 4);
  sc->diffs[4] = score;
  sc->n_diffs++;
  double lo = sc->diffs[0], hi = sc->diffs[0];
  for (int i = 1; i < 4; i++) {
    lo = lo < sc->diffs[i] ? lo : sc->diffs[i];
    hi = hi > sc->diffs[i] ? hi : sc->diffs[i];
  }
  double threshold = 1.8 * hi - 0.8 * lo;
  if (sc->n_diffs > 4) {
    if (score < 5) change = 0;
    else if (score / threshold < 1.0) change = 0;
    else if (score > 30 && score / sc->diffs[3] > 1.4) change = 1;
    else if (score / threshold > 2.3) change = 1;
    else if (score > 50) change = 1;
    else change = 0;
  } else change = 0;
  if (change) { memset (sc->diffs, 0, sizeof sc->diffs); sc->n_diffs = 0; }
  return change;
}

/* smooth: gst/smooth/gstsmooth.c:131-176 in closed form. Iteration y of the reference
 * reads its reference sample from row max(y-1,0) and writes that row, with window rows
 * [max(0,y-fs-1), min(fs+1,h) + min(y+1, max(0,h-fs-1))); row 0 is written twice (the
 * later iteration wins), the last row never. */
EXPORT void
oracle_smooth_plane (uint8_t *dest, const uint8_t *src, int width, int height, int stride, int dstride,
    int tolerance, int filtersize)
{
  int fs = filtersize;
  for (int y = 0; y < height; y++) {
    int r = y > 0 ? y - 1 : 0;
    long long ra = (long long) y - ((long long) fs + 1); if (ra < 0) ra = 0;
    long long lim = (long long) height - ((long long) fs + 1); if (lim < 0) lim = 0;
    long long rb = ((long long) fs + 1 < height ? (long long) fs + 1 : height) + (y + 1 < lim ? y + 1 : lim);
    for (int x = 0; x < width; x++) {
      int ref = src[(size_t) r * stride + x];
      int upper = ref + tolerance, lower = ref - tolerance;
      int num = 1, sum = ref;
      long long c0 = (long long) x - fs; if (c0 < 0) c0 = 0;
      long long c1 = (long long) x + fs + 1; if (c1 > width) c1 = width;
      for (long long wr = ra; wr < rb; wr++)
        for (long long wc = c0; wc < c1; wc++) {
          int akt = src[(size_t) wr * stride + wc];
          if ((lower - akt) * (upper - akt) < 0) { num++; sum += akt; }
        }
      dest[(size_t) r * dstride + x] = (uint8_t) (sum / num);
    }
  }
}

/* videoanalyse: gst/videosignal/gstvideoanalyse.c:206-236 (two passes, integer average) */
EXPORT void
oracle_videoanalyse (const uint8_t *luma, int stride, int width, int height, double *average, double *variance)
{
  uint64_t sum = 0;
  for (int i = 0; i < height; i++)
    for (int j = 0; j < width; j++) sum += luma[(size_t) i * stride + j];
  int avg = (int) (sum / (uint64_t) (width * height));
  *average = sum / (255.0 * width * height);
  sum = 0;
  for (int i = 0; i < height; i++)
    for (int j = 0; j < width; j++) {
      int diff = avg - luma[(size_t) i * stride + j];
      sum += (uint64_t) (diff * diff);
    }
  *variance = sum / (255.0 * 255.0 * width * height);
}
