// colorops.cu - exclusion, dilate (gaudieffects) and coloreffects / chromahold
// (coloreffects plugin) for sm_100a. All are 4-byte -> 4-byte HBM-bound integer
// work (8 B/px): 128-bit coalesced accesses, lane-replicated shared-memory tables
// (entry v of lane l at word v*32+l: conflict-free for arbitrary data), no tensor cores.
#include "common.cuh"
#include "dilate.cuh"
#include "stream.cuh"
#include <stdlib.h>
#include <string.h>

namespace {

struct Table256 { uint32_t w[256]; };

__device__ __forceinline__ void table_fill (uint32_t *tab, const Table256 &t) {
  for (int i = threadIdx.x + threadIdx.y * blockDim.x; i < 256 * 32; i += blockDim.x * blockDim.y) tab[i] = t.w[i >> 5];
  __syncthreads ();
}
constexpr int TAB_SMEM = 256 * 32 * 4;

__device__ __forceinline__ int clamp255 (int v) { return min (max (v, 0), 255); }

// A frame batch as a 2-D grid of 4-pixel groups: x = group in row, y = row, z = frame.
struct Geom {
  uint8_t *data;          // in-place ops; out-of-place ops carry src separately
  const uint8_t *src;
  size_t frame_stride;
  int row_stride, width, height;
};

// ------------------------------------------------------------------ exclusion
// gst/gaudieffects/gstexclusion.c:256-284. Table word for value v:
//   byte0 = blue'(v) = green'(v) = clamp(f - ((f-v)^2/f + v*v/f)), bits 16..31 = (f-v)^2/f
// red' = clamp(f - ((f-r)^2/f + (g*r)/f)) needs the cross term (:269-270): g*r/f by
// multiply-high with M = floor(2^32/f)+1, exact for g*r < 2^16 (error g*r*f/2^32 < 1/f).
struct ExclParams { Table256 t; uint32_t magic; int factor; };

__device__ __forceinline__ uint32_t excl_px (const uint32_t *tl, uint32_t in, uint32_t magic, int f) {
  uint32_t b = in & 0xff, g = (in >> 8) & 0xff, r = (in >> 16) & 0xff;
  uint32_t wb = tl[b << 5], wg = tl[g << 5], wr = tl[r << 5];
  uint32_t gr = g * r;
  uint32_t q = (f == 1) ? gr : __umulhi (gr, magic);
  int r2 = clamp255 (f - (int) ((wr >> 16) + q));
  return (wb & 0xff) | ((wg & 0xff) << 8) | ((uint32_t) r2 << 16);
}

struct ExclOp {                          // stream.cuh operator
  ExclParams p;
  __device__ __forceinline__ void fill (uint32_t *tab) const { table_fill (tab, p.t); }
  __device__ __forceinline__ uint32_t apply (const uint32_t *tl, uint32_t px) const { return excl_px (tl, px, p.magic, p.factor); }
};

__global__ void __launch_bounds__ (512)
exclusion_kernel (const uint4 *__restrict__ src, uint4 *__restrict__ dst, size_t n16,
    const uint32_t *__restrict__ src_tail, uint32_t *__restrict__ dst_tail, int ntail,
    const __grid_constant__ ExclParams p)
{
  extern __shared__ uint32_t tab[];
  table_fill (tab, p.t);
  const uint32_t *tl = tab + (threadIdx.x & 31);
  const size_t stride = (size_t) gridDim.x * blockDim.x;
  constexpr int U = 4;                                     // 128-bit groups in flight per thread
  size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + (U - 1) * stride < n16; i += U * stride) {
    uint4 v[U];
#pragma unroll
    for (int k = 0; k < U; k++) v[k] = ld_stream_v4 (src + i + k * stride);
#pragma unroll
    for (int k = 0; k < U; k++) {
      uint4 o;
      o.x = excl_px (tl, v[k].x, p.magic, p.factor); o.y = excl_px (tl, v[k].y, p.magic, p.factor);
      o.z = excl_px (tl, v[k].z, p.magic, p.factor); o.w = excl_px (tl, v[k].w, p.magic, p.factor);
      st_stream_v4 (dst + i + k * stride, o);
    }
  }
  for (; i < n16; i += stride) {
    uint4 v = ld_stream_v4 (src + i), o;
    o.x = excl_px (tl, v.x, p.magic, p.factor); o.y = excl_px (tl, v.y, p.magic, p.factor);
    o.z = excl_px (tl, v.z, p.magic, p.factor); o.w = excl_px (tl, v.w, p.magic, p.factor);
    st_stream_v4 (dst + i, o);
  }
  if (blockIdx.x == 0 && threadIdx.x < ntail) dst_tail[threadIdx.x] = excl_px (tl, src_tail[threadIdx.x], p.magic, p.factor);
}

// --------------------------------------------------------------------- dilate
// arithmetic: dilate.cuh; TMA-fed fast path: dilate_tma.cu
__device__ __forceinline__ void dil_pick (uint32_t &best, uint32_t &bl, uint32_t cand, uint32_t cl, bool erode) {
  if (erode) dil_pick_t<true> (best, bl, cand, cl); else dil_pick_t<false> (best, bl, cand, cl);
}

constexpr int DIL_ROWS = 16;   // rows one warp marches down

// width % 4 == 0: a lane owns 4 pixels (one 128-bit word) per row, a warp 128 pixels;
// left/right neighbours by shuffle, the row below is the next iteration's row; every
// pixel's luminance is computed once per row it takes part in (own row, and as `down`).
template <bool ERODE>
__global__ void __launch_bounds__ (256)
dilate_kernel (const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, int width, int height,
    size_t frame_stride, const uint8_t *__restrict__ below)
{
  const int lane = threadIdx.x;
  const int x0 = (blockIdx.x * 32 + lane) * 4;
  const int j0 = (blockIdx.y * blockDim.y + threadIdx.y) * DIL_ROWS;
  if (j0 >= height) return;
  const int jend = min (j0 + DIL_ROWS, height);
  const bool active = x0 < width;
  const bool last_word = x0 + 4 >= width;
  const size_t rs = (size_t) width * 4;
  const uint8_t *s = src + (size_t) blockIdx.z * frame_stride;
  uint8_t *d = dst + (size_t) blockIdx.z * frame_stride;
  const uint8_t *bel = below ? below + (size_t) blockIdx.z * rs : nullptr;

  auto load = [&] (int j, uint4 &px, uint32_t &le, uint32_t &re) {
    const uint8_t *rp = (j < height) ? s + (size_t) j * rs : bel;      // row `height` = the row under the shard
    px = make_uint4 (0, 0, 0, 0); le = re = 0;
    if (active) {
      px = ld_stream_v4 (rp + (size_t) x0 * 4);
      if (lane == 0 && x0 > 0) le = ldg_u32 (rp + (size_t) (x0 - 1) * 4);
      if (lane == 31 && !last_word) re = ldg_u32 (rp + (size_t) (x0 + 4) * 4);
    }
  };

  uint4 cur, nxt; uint32_t cle, cre, nle = 0, nre = 0;
  load (j0, cur, cle, cre);
  uint32_t cl[4] = { dil_lum (cur.x), dil_lum (cur.y), dil_lum (cur.z), dil_lum (cur.w) };
  for (int j = j0; j < jend; j++) {
    const bool has_down = (j + 1 < height) || (bel != nullptr);
    if (has_down) load (j + 1, nxt, nle, nre); else nxt = cur;
    const uint32_t p[4] = { cur.x, cur.y, cur.z, cur.w };
    const uint32_t dn[4] = { nxt.x, nxt.y, nxt.z, nxt.w };
    const uint32_t nl[4] = { dil_lum (nxt.x), dil_lum (nxt.y), dil_lum (nxt.z), dil_lum (nxt.w) };
    uint32_t left_in = __shfl_up_sync (0xffffffffu, cur.w, 1);
    uint32_t right_in = __shfl_down_sync (0xffffffffu, cur.x, 1);
    if (lane == 0) left_in = (x0 > 0) ? cle : cur.x;              // left of column 0 is the pixel itself
    if (last_word) right_in = cur.w;                               // right of the last column is itself
    else if (lane == 31) right_in = cre;
    const uint32_t ll = dil_lum (left_in), rl = dil_lum (right_in);
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      uint32_t best = p[k], bl = cl[k];
      dil_pick_t<ERODE> (best, bl, dn[k], nl[k]);                                   // down
      dil_pick_t<ERODE> (best, bl, (k < 3) ? p[k + 1] : right_in, (k < 3) ? cl[k + 1] : rl);   // right
      dil_pick_t<ERODE> (best, bl, (k > 0) ? p[k - 1] : left_in, (k > 0) ? cl[k - 1] : ll);    // left (`up` is dead code)
      o[k] = best;
    }
    if (active) st_stream_v4 (d + (size_t) j * rs + (size_t) x0 * 4, make_uint4 (o[0], o[1], o[2], o[3]));
    cur = nxt; cle = nle; cre = nre;
#pragma unroll
    for (int k = 0; k < 4; k++) cl[k] = nl[k];
  }
}

// any width (rows not 16-byte aligned): one pixel per thread
__global__ void __launch_bounds__ (256)
dilate_scalar_kernel (const uint32_t *__restrict__ src, uint32_t *__restrict__ dst, int width, int height,
    size_t frame_px, int erode, const uint32_t *__restrict__ below)
{
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= width) return;
  const uint32_t *s = src + (size_t) blockIdx.z * frame_px;
  const uint32_t *p = s + (size_t) y * width + x;
  uint32_t best = *p, bl = dil_lum (best);
  uint32_t dn = (y + 1 < height) ? p[width] : (below ? below[(size_t) blockIdx.z * width + x] : *p);
  uint32_t rg = (x + 1 < width) ? p[1] : *p;
  uint32_t lf = (x > 0) ? p[-1] : *p;
  dil_pick (best, bl, dn, dil_lum (dn), erode);
  dil_pick (best, bl, rg, dil_lum (rg), erode);
  dil_pick (best, bl, lf, dil_lum (lf), erode);
  dst[(size_t) blockIdx.z * frame_px + (size_t) y * width + x] = best;
}

// --------------------------------------------------------------- coloreffects
// gst/coloreffects/gstcoloreffects.c:303-359 (rgb) and :361-435 (ayuv). In place.
// Table word for value v (or luma l): the three table bytes already shifted to the
// pixel's R/G/B (Y/U/V) byte positions, so the luma path is one lookup + one merge.
struct ColorParams {
  Table256 t;             // luma:  w[l] = R<<8*or | G<<8*og | B<<8*ob;  per-channel: same with (t[3v],t[3v+1],t[3v+2])
  uint32_t keep_mask;     // bits of the pixel left untouched (alpha / x)
  int sr, sg, sb;         // bit shifts of the R,G,B (Y,U,V) bytes
  int map_luma;
};

__device__ __forceinline__ uint32_t ce_rgb_px (const uint32_t *tl, uint32_t in, const ColorParams &p) {
  uint32_t r = (in >> p.sr) & 0xff, g = (in >> p.sg) & 0xff, b = (in >> p.sb) & 0xff;
  if (p.map_luma) {
    uint32_t luma = (54u * r + 183u * g + 19u * b) >> 8;       // == (((r<<8)*54 + (g<<8)*183 + (b<<8)*19) >> 16), :337-338
    return (in & p.keep_mask) | tl[luma << 5];
  }
  uint32_t wr = tl[r << 5], wg = tl[g << 5], wb = tl[b << 5];
  return (in & p.keep_mask) | (wr & (0xffu << p.sr)) | (wg & (0xffu << p.sg)) | (wb & (0xffu << p.sb));
}

__device__ __forceinline__ int ce_mat (int a0, int a1, int a2, int a3, int x, int y, int z) {
  return (a0 * x + a1 * y + a2 * z + a3) >> 8;                 // APPLY_MATRIX, :299-301
}

__device__ __forceinline__ uint32_t ce_ayuv_px (const uint32_t *tl, uint32_t in, const ColorParams &p) {
  int y = (in >> p.sr) & 0xff, u = (in >> p.sg) & 0xff, v = (in >> p.sb) & 0xff;
  int r, g, b;
  if (p.map_luma) {
    uint32_t w = tl[y << 5];                                   // here the word holds R,G,B in bytes 0,1,2
    r = w & 0xff; g = (w >> 8) & 0xff; b = (w >> 16) & 0xff;
  } else {
    r = clamp255 (ce_mat (298, 0, 409, -57068, y, u, v));     // cog_ycbcr_to_rgb_matrix_8bit_sdtv, :288-292
    g = clamp255 (ce_mat (298, -100, -208, 34707, y, u, v));
    b = clamp255 (ce_mat (298, 516, 0, -70870, y, u, v));
    r = tl[r << 5] & 0xff; g = (tl[g << 5] >> 8) & 0xff; b = (tl[b << 5] >> 16) & 0xff;
  }
  int y2 = clamp255 (ce_mat (66, 129, 25, 4096, r, g, b));    // cog_rgb_to_ycbcr_matrix_8bit_sdtv, :294-298
  int u2 = clamp255 (ce_mat (-38, -74, 112, 32768, r, g, b));
  int v2 = clamp255 (ce_mat (112, -94, -18, 32768, r, g, b));
  return (in & p.keep_mask) | ((uint32_t) y2 << p.sr) | ((uint32_t) u2 << p.sg) | ((uint32_t) v2 << p.sb);
}

template <bool AYUV>
struct ColorOp {                         // stream.cuh operator
  ColorParams p;
  __device__ __forceinline__ void fill (uint32_t *tab) const { table_fill (tab, p.t); }
  __device__ __forceinline__ uint32_t apply (const uint32_t *tl, uint32_t px) const {
    return AYUV ? ce_ayuv_px (tl, px, p) : ce_rgb_px (tl, px, p);
  }
};

// 4-byte pixels; a row is `width` pixels at data + y*row_stride (rows 4-byte aligned).
template <bool AYUV>
__global__ void __launch_bounds__ (256)
coloreffects4_kernel (uint8_t *data, int width, int height, int row_stride, size_t frame_stride,
    const __grid_constant__ ColorParams p)
{
  extern __shared__ uint32_t tab[];
  table_fill (tab, p.t);
  const uint32_t *tl = tab + (threadIdx.x & 31);
  const int groups = (width + 3) / 4;
  for (int y = blockIdx.y; y < height; y += gridDim.y) {
    uint8_t *row = data + (size_t) blockIdx.z * frame_stride + (size_t) y * row_stride;
    const bool vec = (((uintptr_t) row) & 15) == 0;
    const int gstride = gridDim.x * blockDim.x;
    int gx = blockIdx.x * blockDim.x + threadIdx.x;
    constexpr int U = 4;                                   // full 128-bit groups in flight per thread
    if (vec) {
      const int full = width / 4;
      uint4 *r4 = reinterpret_cast<uint4 *> (row);
      for (; gx + (U - 1) * gstride < full; gx += U * gstride) {
        uint4 v[U];
#pragma unroll
        for (int k = 0; k < U; k++) v[k] = ld_na_v4 (r4 + gx + k * gstride);
#pragma unroll
        for (int k = 0; k < U; k++) {
          uint4 o;
          if (AYUV) { o.x = ce_ayuv_px (tl, v[k].x, p); o.y = ce_ayuv_px (tl, v[k].y, p); o.z = ce_ayuv_px (tl, v[k].z, p); o.w = ce_ayuv_px (tl, v[k].w, p); }
          else { o.x = ce_rgb_px (tl, v[k].x, p); o.y = ce_rgb_px (tl, v[k].y, p); o.z = ce_rgb_px (tl, v[k].z, p); o.w = ce_rgb_px (tl, v[k].w, p); }
          st_stream_v4 (r4 + gx + k * gstride, o);
        }
      }
    }
    for (; gx < groups; gx += gstride) {
      int x0 = gx * 4, n = min (4, width - x0);
      uint32_t *px = reinterpret_cast<uint32_t *> (row) + x0;
      if (vec && n == 4) {
        uint4 v = *reinterpret_cast<uint4 *> (px), o;
        if (AYUV) { o.x = ce_ayuv_px (tl, v.x, p); o.y = ce_ayuv_px (tl, v.y, p); o.z = ce_ayuv_px (tl, v.z, p); o.w = ce_ayuv_px (tl, v.w, p); }
        else { o.x = ce_rgb_px (tl, v.x, p); o.y = ce_rgb_px (tl, v.y, p); o.z = ce_rgb_px (tl, v.z, p); o.w = ce_rgb_px (tl, v.w, p); }
        *reinterpret_cast<uint4 *> (px) = o;
      } else {
        for (int k = 0; k < n; k++) px[k] = AYUV ? ce_ayuv_px (tl, px[k], p) : ce_rgb_px (tl, px[k], p);
      }
    }
  }
}

// 3-byte pixels (RGB / BGR): 4 pixels = 12 bytes = three aligned words per thread.
// Table word: bytes 0,1,2 = table[3v], table[3v+1], table[3v+2].
struct Color3Params { Table256 t; int o_r, o_g, o_b; int map_luma; };

__device__ __forceinline__ void ce_rgb3 (const uint32_t *tl, uint8_t *px, const Color3Params &p) {
  uint32_t r = px[p.o_r], g = px[p.o_g], b = px[p.o_b];
  if (p.map_luma) {
    uint32_t w = tl[((54u * r + 183u * g + 19u * b) >> 8) << 5];
    px[p.o_r] = w & 0xff; px[p.o_g] = (w >> 8) & 0xff; px[p.o_b] = (w >> 16) & 0xff;
  } else {
    px[p.o_r] = tl[r << 5] & 0xff; px[p.o_g] = (tl[g << 5] >> 8) & 0xff; px[p.o_b] = (tl[b << 5] >> 16) & 0xff;
  }
}

__global__ void __launch_bounds__ (256)
coloreffects3_kernel (uint8_t *data, int width, int height, int row_stride, size_t frame_stride,
    const __grid_constant__ Color3Params p)
{
  extern __shared__ uint32_t tab[];
  table_fill (tab, p.t);
  const uint32_t *tl = tab + (threadIdx.x & 31);
  const int groups = (width + 3) / 4;
  for (int y = blockIdx.y; y < height; y += gridDim.y) {
    uint8_t *row = data + (size_t) blockIdx.z * frame_stride + (size_t) y * row_stride;
    const bool al = (((uintptr_t) row) & 3) == 0;
    for (int gx = blockIdx.x * blockDim.x + threadIdx.x; gx < groups; gx += gridDim.x * blockDim.x) {
      int x0 = gx * 4, n = min (4, width - x0);
      uint8_t *px = row + (size_t) x0 * 3;
      if (al && n == 4) {
        uint32_t w[3];
        uint32_t *wp = reinterpret_cast<uint32_t *> (px);
        w[0] = wp[0]; w[1] = wp[1]; w[2] = wp[2];
        uint8_t *bytes = reinterpret_cast<uint8_t *> (w);
#pragma unroll
        for (int k = 0; k < 4; k++) ce_rgb3 (tl, bytes + 3 * k, p);
        wp[0] = w[0]; wp[1] = w[1]; wp[2] = w[2];
      } else {
        for (int k = 0; k < n; k++) ce_rgb3 (tl, px + 3 * k, p);
      }
    }
  }
}

// ----------------------------------------------------------------- chromahold
// gst/coloreffects/gstchromahold.c:271-360. Table: w[C] = floor(2^32/C)+1 (C >= 2):
// |256*60*d + C/2| < 2^22 and C <= 255, so multiply-high gives the exact quotient.
struct ChromaParams {
  Table256 t; int sr, sg, sb; int h1; int tolerance; uint32_t keep_mask;
  uint32_t rep;            // 1 at the bit positions of the three channels: grey * rep replicates the grey level into them
  uint32_t w_lo, w_hi;     // grey weights of the bytes (0, 1) and (2, 3) of a pixel word as 16-bit pairs (0 for the x / alpha byte)
  int branchy;             // A/B knob: the literal control flow instead of selects
  uint32_t sel_r, sel_g, sel_b;   // PRMT selectors: channel byte, zero-extended
  int tol_eff;             // tolerance, or -1 when h1 == -1 (every pixel goes grey)
};

// The literal control flow of rgb_to_hue / hue_dist / the process loop (A/B knob B200VF_CHROMA_BRANCHY=1): shorter per
// path, but every data-dependent branch splits the warp on frames without spatial coherence.
__device__ __forceinline__ int ch_div (int num, int C, const uint32_t *tl) {
  uint32_t a = (uint32_t) abs (num);
  uint32_t q = (C == 1) ? a : __umulhi (a, tl[C << 5]);
  return num < 0 ? -(int) q : (int) q;                     // C division truncates toward zero
}
__device__ __forceinline__ int ch_hue_branchy (int r, int g, int b, const uint32_t *tl) {
  int m = min (min (r, g), b), M = max (max (r, g), b);
  int C = M - m, C2 = C >> 1, h;
  if (C == 0) return -1;                                   // G_MAXUINT as gint (:282)
  if (M == r) h = ch_div (256 * 60 * (g - b) + C2, C, tl);
  else if (M == g) h = ch_div (256 * 60 * (b - r) + C2, C, tl) + 120 * 256;
  else h = ch_div (256 * 60 * (r - g) + C2, C, tl) + 240 * 256;
  h >>= 8;
  if (h >= 360) h -= 360; else if (h < 0) h += 360;
  return h;
}
__device__ __forceinline__ uint32_t ch_px_branchy (const uint32_t *tl, uint32_t in, const ChromaParams &p) {
  int r = (in >> p.sr) & 0xff, g = (in >> p.sg) & 0xff, b = (in >> p.sb) & 0xff;
  int h2 = ch_hue_branchy (r, g, b, tl);
  int d1 = p.h1 - h2, d2 = h2 - p.h1;
  if (d1 < 0) d1 += 360;
  if (d2 < 0) d2 += 360;
  int diff = min (d1, d2);
  if (p.h1 == -1 || diff > p.tolerance) {
    const uint32_t grey = __dp2a_hi (p.w_hi, in, __dp2a_lo (p.w_lo, in, 0u)) >> 16;
    return (in & p.keep_mask) + grey * p.rep;
  }
  return in;
}

// Branch-free: the three hue sectors, the grey pixels (C == 0) and the hold / grey decision depend on the data, and on
// random frames every warp diverged at each of them (BSSY / BSYNC around every branch: 4 per pixel in the SASS); selects
// cost a few instructions more per path and far fewer per warp.
// Returns the hue WITHOUT the reference's wrap into [0, 360) (:292-296): -60 .. 299, or -1 for grey pixels. The only
// consumer is the circular distance below, which does not care (checked over all 2^24 colours: tools note in
// profiles/r02_chromahold.md). The three right shifts go through multiply-high: the ALU pipe (half rate) bounds this
// kernel, the FMA pipe has room.
__device__ __forceinline__ int ch_hue (int r, int g, int b, const uint32_t *tl) {
  const int m = min (min (r, g), b), M = max (max (r, g), b);
  const int C = M - m;
  const bool ir = (M == r), ig = (M == g);
  const int x = ir ? g : (ig ? b : r), y = ir ? b : (ig ? r : g);
  const int off = ir ? 0 : (ig ? 120 * 256 : 240 * 256);
  const int num = 256 * 60 * (x - y) + (int) __umulhi ((uint32_t) C, 0x80000000u);      // + (C >> 1)
  const uint32_t a = (uint32_t) abs (num);
  uint32_t q = __umulhi (a, tl[C << 5]);                   // exact quotient for C >= 2 (table comment above)
  q = (C == 1) ? a : q;
  const int h = ((num < 0) ? -(int) q : (int) q) + off;    // C division truncates toward zero
  return (C == 0) ? -1 : __mulhi (h, 1 << 24);             // h >> 8 (arithmetic); G_MAXUINT as gint for greys (:282)
}

__device__ __forceinline__ uint32_t ch_px (const uint32_t *tl, uint32_t in, const ChromaParams &p) {
  // channel bytes by PRMT (one ALU instruction each; shift + mask are two)
  const int r = (int) __byte_perm (in, 0u, p.sel_r), g = (int) __byte_perm (in, 0u, p.sel_g), b = (int) __byte_perm (in, 0u, p.sel_b);
  const int h2 = ch_hue (r, g, b, tl);
  // hue_dist (:301-315) = the circular distance of the two hues: d1 = h1 - h2, d2 = h2 - h1, each + 360 when negative,
  // the smaller one. With a = |h1 - h2| and h2 possibly unwrapped (a <= 420) that is min (a, |360 - a|) - also for
  // h2 = -1 (grey pixel): min (h1 + 1, 359 - h1).
  const int a = abs (p.h1 - h2);
  const int diff = min (a, abs (360 - a));
  // (13938 r + 46869 g + 4730 b) >> 16 (:345-347) by two 16 x 8-bit dot products on the word itself; the weights sum
  // to 65537, so the result is at most 255 * 65537 >> 16 = 255 and the reference's CLAMP never acts
  const uint32_t grey = __umulhi (__dp2a_hi (p.w_hi, in, __dp2a_lo (p.w_lo, in, 0u)), 65536u);
  const uint32_t greyed = (in & p.keep_mask) + grey * p.rep;
  return (diff > p.tol_eff) ? greyed : in;                 // tol_eff = -1 when the target itself is grey (h1 == -1: always)
}

struct ChromaOp {                        // stream.cuh operator
  ChromaParams p;
  __device__ __forceinline__ void fill (uint32_t *tab) const { table_fill (tab, p.t); }
  __device__ __forceinline__ uint32_t apply (const uint32_t *tl, uint32_t px) const {
    return p.branchy ? ch_px_branchy (tl, px, p) : ch_px (tl, px, p);
  }
};

__global__ void __launch_bounds__ (256)
chromahold_kernel (uint8_t *data, int width, int height, int row_stride, size_t frame_stride,
    const __grid_constant__ ChromaParams p)
{
  extern __shared__ uint32_t tab[];
  table_fill (tab, p.t);
  const uint32_t *tl = tab + (threadIdx.x & 31);
  const int groups = (width + 3) / 4;
  for (int y = blockIdx.y; y < height; y += gridDim.y) {
    uint8_t *row = data + (size_t) blockIdx.z * frame_stride + (size_t) y * row_stride;
    const bool vec = (((uintptr_t) row) & 15) == 0;
    const int gstride = gridDim.x * blockDim.x;
    int gx = blockIdx.x * blockDim.x + threadIdx.x;
    constexpr int U = 4;
    if (vec) {
      const int full = width / 4;
      uint4 *r4 = reinterpret_cast<uint4 *> (row);
      for (; gx + (U - 1) * gstride < full; gx += U * gstride) {
        uint4 v[U];
#pragma unroll
        for (int k = 0; k < U; k++) v[k] = ld_na_v4 (r4 + gx + k * gstride);
#pragma unroll
        for (int k = 0; k < U; k++) {
          uint4 o;
          o.x = ch_px (tl, v[k].x, p); o.y = ch_px (tl, v[k].y, p); o.z = ch_px (tl, v[k].z, p); o.w = ch_px (tl, v[k].w, p);
          st_stream_v4 (r4 + gx + k * gstride, o);
        }
      }
    }
    for (; gx < groups; gx += gstride) {
      int x0 = gx * 4, n = min (4, width - x0);
      uint32_t *px = reinterpret_cast<uint32_t *> (row) + x0;
      if (vec && n == 4) {
        uint4 v = *reinterpret_cast<uint4 *> (px), o;
        o.x = ch_px (tl, v.x, p); o.y = ch_px (tl, v.y, p); o.z = ch_px (tl, v.z, p); o.w = ch_px (tl, v.w, p);
        *reinterpret_cast<uint4 *> (px) = o;
      } else {
        for (int k = 0; k < n; k++) px[k] = ch_px (tl, px[k], p);
      }
    }
  }
}

// rows of one frame batch -> a 2-D persistent-ish grid: x covers a row's 4-pixel groups,
// y strides over rows; total CTAs a small multiple of the SM count.
// A contiguous batch (row_stride == 4*width, frames back to back) is flattened to
// one long row so every access is a full 128-bit vector.
struct Launch2D { dim3 grid, block; int width, height, row_stride; size_t frame_stride; int nframes; };

// contiguous frames of 4-byte pixels whose bytes form one 16-byte aligned stream: the TMA ring takes them
bool flat_stream (const b200vf_ctx *ctx, const uint8_t *data, int width, int height, int row_stride, size_t frame_stride,
    int nframes, size_t *nbytes) {
  if (!stream_enabled (ctx) || row_stride != 4 * width) return false;
  if (nframes > 1 && frame_stride != (size_t) row_stride * height) return false;
  const size_t n = (size_t) row_stride * height * nframes;
  if (((uintptr_t) data) % 16 != 0 || n % 16 != 0 || n < 16384) return false;
  *nbytes = n;
  return true;
}

Launch2D plan2d (b200vf_ctx *ctx, int width, int height, int row_stride, size_t frame_stride, int nframes, int pstride) {
  Launch2D l;
  l.width = width; l.height = height; l.row_stride = row_stride; l.frame_stride = frame_stride; l.nframes = nframes;
  if (row_stride == pstride * width && (size_t) width * height < (1u << 30) / 4) {
    l.width = width * height; l.height = 1; l.row_stride = pstride * width * height;
    if (frame_stride == (size_t) row_stride * height && (size_t) l.width * nframes < (1u << 30) / 4) {
      l.width *= nframes; l.nframes = 1; l.frame_stride = 0;
    }
  }
  l.block = dim3 (256, 1, 1);
  int groups = (l.width + 3) / 4;
  int gx = (groups + 255) / 256;
  int per_sm = 4;                                 // CTAs of 256 threads per SM (32 KB table each)
  if (const char *e = getenv ("B200VF_PLAN2D_CTAS_PER_SM")) { int v = atoi (e); if (v >= 1 && v <= 7) per_sm = v; }   // tuning knob
  int target = ctx->sm_count * per_sm;
  int gy = 1;
  if (gx > target) gx = target;
  else { gy = target / gx; if (gy > l.height) gy = l.height; if (gy < 1) gy = 1; }
  if (gy > 65535) gy = 65535;
  l.grid = dim3 (gx, gy, l.nframes);
  return l;
}

int shift_of (int off) { return 8 * off; }

}  // namespace

// ------------------------------------------------------------------- C entries
B200VF_API int b200vf_exclusion (b200vf_ctx *ctx, const uint8_t *d_src, uint8_t *d_dst, size_t npix_total,
    int factor, void *stream)
{
  B200VF_REQUIRE (ctx && d_src && d_dst, B200VF_E_INVAL, "exclusion: NULL argument");
  B200VF_REQUIRE (factor >= 1 && factor <= 175, B200VF_E_PROPERTY, "exclusion: factor %d not in [1,175]", factor);
  B200VF_REQUIRE (((uintptr_t) d_src) % 16 == 0 && ((uintptr_t) d_dst) % 16 == 0, B200VF_E_INVAL,
      "exclusion: buffers must be 16-byte aligned");
  if (!npix_total) return B200VF_OK;
  ExclParams p;
  for (int v = 0; v < 256; v++) {
    int sq = (factor - v) * (factor - v) / factor;
    int gb = factor - (sq + (v * v) / factor);
    gb = gb > 255 ? 255 : (gb < 0 ? 0 : gb);
    p.t.w[v] = (uint32_t) gb | ((uint32_t) sq << 16);
  }
  p.magic = factor >= 2 ? (uint32_t) (0x100000000ull / (uint64_t) factor) + 1u : 0u;
  p.factor = factor;
  if (int rc = b200vf_func_smem (ctx, (const void *) exclusion_kernel, TAB_SMEM)) return rc;
  size_t n16 = npix_total / 4;
  int ntail = (int) (npix_total - n16 * 4);
  if (stream_enabled (ctx) && n16 >= 1024) {
    ExclOp op;
    op.p = p;
    int rc = stream_launch (ctx, d_src, d_dst, n16 * 16, op, b200vf_stream (ctx, stream), "exclusion_tma");
    if (rc || !ntail) return rc;
    exclusion_kernel<<<1, 512, TAB_SMEM, b200vf_stream (ctx, stream)>>> (nullptr, nullptr, 0,
        reinterpret_cast<const uint32_t *> (d_src) + n16 * 4, reinterpret_cast<uint32_t *> (d_dst) + n16 * 4, ntail, p);
    return b200vf_launched (ctx, "exclusion_tail");
  }
  int grid = ctx->sm_count * 2;
  size_t need = (n16 + 511) / 512;
  if (need < (size_t) grid) grid = need ? (int) need : 1;
  exclusion_kernel<<<grid, 512, TAB_SMEM, b200vf_stream (ctx, stream)>>> (reinterpret_cast<const uint4 *> (d_src),
      reinterpret_cast<uint4 *> (d_dst), n16, reinterpret_cast<const uint32_t *> (d_src) + n16 * 4,
      reinterpret_cast<uint32_t *> (d_dst) + n16 * 4, ntail, p);
  return b200vf_launched (ctx, "exclusion");
}

B200VF_API int b200vf_dilate (b200vf_ctx *ctx, const uint8_t *d_src, uint8_t *d_dst, int width, int height,
    size_t frame_stride, int nframes, int erode, const uint8_t *d_below, void *stream)
{
  B200VF_REQUIRE (ctx && d_src && d_dst && width > 0 && height > 0 && nframes > 0, B200VF_E_INVAL, "dilate: bad argument");
  B200VF_REQUIRE (frame_stride >= (size_t) width * height * 4 && frame_stride % 4 == 0 &&
      ((uintptr_t) d_src) % 4 == 0 && ((uintptr_t) d_dst) % 4 == 0, B200VF_E_INVAL, "dilate: frame stride / alignment");
  cudaStream_t s = b200vf_stream (ctx, stream);
  bool vec = (width % 4 == 0) && ((uintptr_t) d_src) % 16 == 0 && ((uintptr_t) d_dst) % 16 == 0 &&
      frame_stride % 16 == 0 && (!d_below || ((uintptr_t) d_below) % 16 == 0);
  const bool tma = vec && frame_stride % 16 == 0 && ctx->variant != 1 && !getenv ("B200VF_DILATE_NO_TMA");
  if (tma) {
    // rows with a row under them inside d_src; a shard's last row takes `down` from d_below: direct kernel, 1 row
    const int rows_tma = d_below ? height - 1 : height;
    if (rows_tma > 0) {
      int rc = b200vf_dilate_tma (ctx, d_src, d_dst, width, height, rows_tma, frame_stride, nframes, erode, s);
      if (rc) return rc;
    }
    if (!d_below) return B200VF_OK;
    const size_t last = (size_t) (height - 1) * width * 4;
    dim3 block (32, 8);
    dim3 grid ((width + 127) / 128, 1, nframes);
    if (erode) dilate_kernel<true><<<grid, block, 0, s>>> (d_src + last, d_dst + last, width, 1, frame_stride, d_below);
    else dilate_kernel<false><<<grid, block, 0, s>>> (d_src + last, d_dst + last, width, 1, frame_stride, d_below);
    return b200vf_launched (ctx, "dilate");
  }
  if (vec) {
    dim3 block (32, 8);
    int strips = (height + DIL_ROWS - 1) / DIL_ROWS;
    dim3 grid ((width + 127) / 128, (strips + 7) / 8, nframes);
    int pad = 0;                                  // tuning knob: unused dynamic shared memory caps the resident CTAs per SM
    if (const char *e = getenv ("B200VF_DILATE_SMEM_PAD_KB")) { int v = atoi (e); if (v >= 0 && v <= 200) pad = v * 1024; }
    if (pad) {
      B200VF_CHECK_CUDA (cudaFuncSetAttribute (dilate_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, pad));
      B200VF_CHECK_CUDA (cudaFuncSetAttribute (dilate_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, pad));
    }
    if (erode) dilate_kernel<true><<<grid, block, pad, s>>> (d_src, d_dst, width, height, frame_stride, d_below);
    else dilate_kernel<false><<<grid, block, pad, s>>> (d_src, d_dst, width, height, frame_stride, d_below);
    return b200vf_launched (ctx, "dilate");
  }
  dim3 grid ((width + 255) / 256, height, nframes);
  dilate_scalar_kernel<<<grid, 256, 0, s>>> (reinterpret_cast<const uint32_t *> (d_src), reinterpret_cast<uint32_t *> (d_dst),
      width, height, frame_stride / 4, erode, reinterpret_cast<const uint32_t *> (d_below));
  return b200vf_launched (ctx, "dilate_scalar");
}

B200VF_API int b200vf_coloreffects_rgb (b200vf_ctx *ctx, uint8_t *d_data, int width, int height, int row_stride,
    size_t frame_stride, int nframes, int pixel_stride, int off_r, int off_g, int off_b,
    const uint8_t *table768, int map_luma, void *stream)
{
  B200VF_REQUIRE (ctx && d_data && width > 0 && height > 0 && nframes > 0, B200VF_E_INVAL, "coloreffects: bad argument");
  if (!table768) return B200VF_OK;                          // preset none: no-op (gstcoloreffects.c:488-490)
  B200VF_REQUIRE (pixel_stride == 3 || pixel_stride == 4, B200VF_E_UNSUPPORTED, "coloreffects: pixel stride %d", pixel_stride);
  B200VF_REQUIRE (row_stride >= pixel_stride * width, B200VF_E_INVAL, "coloreffects: row stride");
  int mo = pixel_stride - 1;
  B200VF_REQUIRE (off_r >= 0 && off_r <= mo && off_g >= 0 && off_g <= mo && off_b >= 0 && off_b <= mo &&
      off_r != off_g && off_g != off_b && off_r != off_b, B200VF_E_INVAL, "coloreffects: offsets");
  cudaStream_t s = b200vf_stream (ctx, stream);
  {
    int rc;
    if ((rc = b200vf_func_smem (ctx, (const void *) coloreffects4_kernel<false>, TAB_SMEM)) ||
        (rc = b200vf_func_smem (ctx, (const void *) coloreffects4_kernel<true>, TAB_SMEM)) ||
        (rc = b200vf_func_smem (ctx, (const void *) coloreffects3_kernel, TAB_SMEM))) return rc;
  }
  if (pixel_stride == 4) {
    B200VF_REQUIRE (((uintptr_t) d_data) % 4 == 0 && row_stride % 4 == 0 && frame_stride % 4 == 0, B200VF_E_INVAL,
        "coloreffects: 4-byte pixels must be 4-byte aligned");
    ColorParams p;
    p.sr = shift_of (off_r); p.sg = shift_of (off_g); p.sb = shift_of (off_b);
    p.keep_mask = ~((0xffu << p.sr) | (0xffu << p.sg) | (0xffu << p.sb));
    p.map_luma = map_luma;
    for (int v = 0; v < 256; v++)
      p.t.w[v] = ((uint32_t) table768[3 * v] << p.sr) | ((uint32_t) table768[3 * v + 1] << p.sg) | ((uint32_t) table768[3 * v + 2] << p.sb);
    size_t nbytes;
    if (flat_stream (ctx, d_data, width, height, row_stride, frame_stride, nframes, &nbytes)) {
      ColorOp<false> op;
      op.p = p;
      return stream_launch (ctx, d_data, d_data, nbytes, op, s, "coloreffects_rgb4_tma", 16);
    }
    Launch2D l = plan2d (ctx, width, height, row_stride, frame_stride, nframes, 4);
    coloreffects4_kernel<false><<<l.grid, l.block, TAB_SMEM, s>>> (d_data, l.width, l.height, l.row_stride, l.frame_stride, p);
    return b200vf_launched (ctx, "coloreffects_rgb4");
  }
  Color3Params p;
  p.o_r = off_r; p.o_g = off_g; p.o_b = off_b; p.map_luma = map_luma;
  for (int v = 0; v < 256; v++)
    p.t.w[v] = (uint32_t) table768[3 * v] | ((uint32_t) table768[3 * v + 1] << 8) | ((uint32_t) table768[3 * v + 2] << 16);
  Launch2D l = plan2d (ctx, width, height, row_stride, frame_stride, nframes, 3);
  if (l.height == 1 && height > 1) { /* flattened only when rows are contiguous */ }
  coloreffects3_kernel<<<l.grid, l.block, TAB_SMEM, s>>> (d_data, l.width, l.height, l.row_stride, l.frame_stride, p);
  return b200vf_launched (ctx, "coloreffects_rgb3");
}

B200VF_API int b200vf_coloreffects_ayuv (b200vf_ctx *ctx, uint8_t *d_data, int width, int height, int row_stride,
    size_t frame_stride, int nframes, int off_y, int off_u, int off_v,
    const uint8_t *table768, int map_luma, void *stream)
{
  B200VF_REQUIRE (ctx && d_data && width > 0 && height > 0 && nframes > 0, B200VF_E_INVAL, "coloreffects_ayuv: bad argument");
  if (!table768) return B200VF_OK;
  B200VF_REQUIRE (row_stride >= 4 * width && ((uintptr_t) d_data) % 4 == 0 && row_stride % 4 == 0 && frame_stride % 4 == 0,
      B200VF_E_INVAL, "coloreffects_ayuv: stride / alignment");
  B200VF_REQUIRE (off_y >= 0 && off_y <= 3 && off_u >= 0 && off_u <= 3 && off_v >= 0 && off_v <= 3 &&
      off_y != off_u && off_u != off_v && off_y != off_v, B200VF_E_INVAL, "coloreffects_ayuv: offsets");
  if (int rc = b200vf_func_smem (ctx, (const void *) coloreffects4_kernel<true>, TAB_SMEM)) return rc;
  ColorParams p;
  p.sr = shift_of (off_y); p.sg = shift_of (off_u); p.sb = shift_of (off_v);
  p.keep_mask = ~((0xffu << p.sr) | (0xffu << p.sg) | (0xffu << p.sb));
  p.map_luma = map_luma;
  for (int v = 0; v < 256; v++)      // R,G,B of the table in bytes 0,1,2 (the matrices need them as numbers)
    p.t.w[v] = (uint32_t) table768[3 * v] | ((uint32_t) table768[3 * v + 1] << 8) | ((uint32_t) table768[3 * v + 2] << 16);
  size_t nbytes;
  if (flat_stream (ctx, d_data, width, height, row_stride, frame_stride, nframes, &nbytes)) {
    ColorOp<true> op;
    op.p = p;
    return stream_launch (ctx, d_data, d_data, nbytes, op, b200vf_stream (ctx, stream), "coloreffects_ayuv_tma", 16);
  }
  Launch2D l = plan2d (ctx, width, height, row_stride, frame_stride, nframes, 4);
  coloreffects4_kernel<true><<<l.grid, l.block, TAB_SMEM, b200vf_stream (ctx, stream)>>> (d_data, l.width, l.height,
      l.row_stride, l.frame_stride, p);
  return b200vf_launched (ctx, "coloreffects_ayuv");
}

B200VF_API int b200vf_chromahold (b200vf_ctx *ctx, uint8_t *d_data, int width, int height, int row_stride,
    size_t frame_stride, int nframes, int off_r, int off_g, int off_b,
    int target_r, int target_g, int target_b, int tolerance, void *stream)
{
  B200VF_REQUIRE (ctx && d_data && width > 0 && height > 0 && nframes > 0, B200VF_E_INVAL, "chromahold: bad argument");
  B200VF_REQUIRE (row_stride >= 4 * width && ((uintptr_t) d_data) % 4 == 0 && row_stride % 4 == 0 && frame_stride % 4 == 0,
      B200VF_E_INVAL, "chromahold: stride / alignment");
  B200VF_REQUIRE (off_r >= 0 && off_r <= 3 && off_g >= 0 && off_g <= 3 && off_b >= 0 && off_b <= 3 &&
      off_r != off_g && off_g != off_b && off_r != off_b, B200VF_E_INVAL, "chromahold: offsets");
  B200VF_REQUIRE (target_r >= 0 && target_r <= 255 && target_g >= 0 && target_g <= 255 && target_b >= 0 && target_b <= 255 &&
      tolerance >= 0 && tolerance <= 180, B200VF_E_PROPERTY, "chromahold: target/tolerance out of range");
  if (int rc = b200vf_func_smem (ctx, (const void *) chromahold_kernel, TAB_SMEM)) return rc;
  ChromaParams p;
  p.sr = shift_of (off_r); p.sg = shift_of (off_g); p.sb = shift_of (off_b);
  p.keep_mask = ~((0xffu << p.sr) | (0xffu << p.sg) | (0xffu << p.sb));
  p.rep = (1u << p.sr) | (1u << p.sg) | (1u << p.sb);
  {
    uint32_t wb[4] = { 0, 0, 0, 0 };                    // weight of byte i of the pixel word
    wb[p.sr / 8] = 13938; wb[p.sg / 8] = 46869; wb[p.sb / 8] = 4730;
    p.w_lo = wb[0] | (wb[1] << 16);
    p.w_hi = wb[2] | (wb[3] << 16);
  }
  p.tolerance = tolerance;
  p.branchy = getenv ("B200VF_CHROMA_BRANCHY") ? 1 : 0;
  p.sel_r = 0x4440u | (unsigned) (p.sr / 8); p.sel_g = 0x4440u | (unsigned) (p.sg / 8); p.sel_b = 0x4440u | (unsigned) (p.sb / 8);
  {   // rgb_to_hue of the target on the host (init_params, gstchromahold.c:362-366)
    int r = target_r, g = target_g, b = target_b;
    int m = r < g ? r : g; if (b < m) m = b;
    int M = r > g ? r : g; if (b > M) M = b;
    int C = M - m, C2 = C >> 1, h;
    if (C == 0) h = -1;
    else {
      if (M == r) h = ((256 * 60 * (g - b) + C2) / C);
      else if (M == g) h = ((256 * 60 * (b - r) + C2) / C) + 120 * 256;
      else h = ((256 * 60 * (r - g) + C2) / C) + 240 * 256;
      h >>= 8;
      if (h >= 360) h -= 360; else if (h < 0) h += 360;
    }
    p.h1 = h;
  }
  p.tol_eff = (p.h1 == -1) ? -1 : tolerance;
  p.t.w[0] = p.t.w[1] = 0;
  for (int c = 2; c < 256; c++) p.t.w[c] = (uint32_t) (0x100000000ull / (uint64_t) c) + 1u;
  size_t nbytes;
  if (flat_stream (ctx, d_data, width, height, row_stride, frame_stride, nframes, &nbytes)) {
    ChromaOp op;
    op.p = p;
    return stream_launch (ctx, d_data, d_data, nbytes, op, b200vf_stream (ctx, stream), "chromahold_tma", 16);
  }
  Launch2D l = plan2d (ctx, width, height, row_stride, frame_stride, nframes, 4);
  chromahold_kernel<<<l.grid, l.block, TAB_SMEM, b200vf_stream (ctx, stream)>>> (d_data, l.width, l.height,
      l.row_stride, l.frame_stride, p);
  return b200vf_launched (ctx, "chromahold");
}
