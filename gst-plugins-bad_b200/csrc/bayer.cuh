// bayer.cuh - the Bayer demosaic algebra on packed bytes, shared by the direct
// and the TMA kernels (and the fused chain).
//
// Closed form of gst/bayer/gstbayer2rgb.c:354-451 + gstbayerorc.orc:3-248
// (SURVEY.md §8a3/a4), written for the BGGR arrangement like the reference:
// a 32-bit word holds 4 horizontally consecutive samples / channel values.
#pragma once
#include "common.cuh"
#include "lut.cuh"

struct BayerRow {     // the two horizontally upsampled lines of one Bayer row
  uint32_t h0;        // even-column samples, odd columns interpolated
  uint32_t h1;        // odd-column samples, even columns interpolated
};

// s = samples x0..x0+3, prev/next = the words left/right of it.
// selL/selR: PRMT selectors patching the frame edges (0x3210 = interior):
//   left edge  (x0 == 0):     s[-1] := s[1]                    (gstbayer2rgb.c:360-363)
//   right edge (last word):   columns n-2,n-1 copy, never average (:372-380)
__device__ __forceinline__ BayerRow bayer_upsample (uint32_t prev, uint32_t s, uint32_t next,
    uint32_t selL, uint32_t selR) {
  uint32_t L = PRMT (prev, s, 0x6543);          // s[x-1] for each of the 4 bytes
  uint32_t R = PRMT (s, next, 0x4321);          // s[x+1]
  L = PRMT (L, R, selL);
  R = PRMT (R, L, selR);
  uint32_t A = avg4 (L, R);                     // avgub(s[x-1], s[x+1])
  BayerRow r;
  r.h0 = PRMT (s, A, 0x7250);                   // [s0, A1, s2, A3]
  r.h1 = PRMT (s, A, 0x3614);                   // [A0, s1, A2, s3]
  return r;
}

// interior words: no edge patching
__device__ __forceinline__ BayerRow bayer_upsample_interior (uint32_t prev, uint32_t s, uint32_t next) {
  uint32_t L = PRMT (prev, s, 0x6543);
  uint32_t R = PRMT (s, next, 0x4321);
  uint32_t A = avg4 (L, R);
  BayerRow r;
  r.h0 = PRMT (s, A, 0x7250);
  r.h1 = PRMT (s, A, 0x3614);
  return r;
}

__device__ __forceinline__ uint32_t bayer_selL (int x0) { return x0 == 0 ? 0x3214u : 0x3210u; }
// v = number of valid pixels in this word (2 or 4); last = this word holds columns n-2,n-1
__device__ __forceinline__ uint32_t bayer_selR (bool last, int v) {
  return !last ? 0x3210u : (v == 4 ? 0x7610u : 0x3254u);
}

// One output row from the upsampled rows above (u), at (c) and below (d).
// gr_row = false: bayer_orc_merge_bg_* (B at even x); true: bayer_orc_merge_gr_*.
__device__ __forceinline__ void bayer_merge (const BayerRow &u, const BayerRow &c, const BayerRow &d,
    bool gr_row, uint32_t &R, uint32_t &G, uint32_t &B) {
  uint32_t va0 = avg4 (u.h0, d.h0);
  uint32_t va1 = avg4 (u.h1, d.h1);
  if (!gr_row) {
    B = c.h0;
    R = va1;
    G = PRMT (avg4 (va0, c.h1), c.h1, 0x7250);  // even x: averaged, odd x: the real sample
  } else {
    R = c.h1;
    B = va0;
    G = PRMT (c.h0, avg4 (va1, c.h0), 0x7250);  // even x: the real sample, odd x: averaged
  }
}

// the same with the row's role known at compile time (unrolled strips)
template <bool GR_ROW>
__device__ __forceinline__ void bayer_merge_ct (const BayerRow &u, const BayerRow &c, const BayerRow &d,
    uint32_t &R, uint32_t &G, uint32_t &B) {
  const uint32_t va0 = avg4 (u.h0, d.h0);
  const uint32_t va1 = avg4 (u.h1, d.h1);
  if (!GR_ROW) {
    B = c.h0;
    R = va1;
    G = PRMT (avg4 (va0, c.h1), c.h1, 0x7250);
  } else {
    R = c.h1;
    B = va0;
    G = PRMT (c.h0, avg4 (va1, c.h0), 0x7250);
  }
}

// 4 channel words (byte position 0..3 of the output pixel) -> 4 packed pixels
__device__ __forceinline__ uint4 bayer_interleave (uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) {
  uint32_t tlo = PRMT (c0, c1, 0x5140), ulo = PRMT (c2, c3, 0x5140);
  uint32_t thi = PRMT (c0, c1, 0x7362), uhi = PRMT (c2, c3, 0x7362);
  uint4 o;
  o.x = PRMT (tlo, ulo, 0x5410);
  o.y = PRMT (tlo, ulo, 0x7632);
  o.z = PRMT (thi, uhi, 0x5410);
  o.w = PRMT (thi, uhi, 0x7632);
  return o;
}

// ORDER: 0 = (r,g,b) at bytes (0,1,2) "rgba", 1 = (2,1,0) "bgra",
//        2 = (1,2,3) "argb",               3 = (3,2,1) "abgr"   (gstbayer2rgb.c:409-421)
template <int ORDER>
__device__ __forceinline__ uint4 bayer_pack (uint32_t R, uint32_t G, uint32_t B, uint32_t A) {
  if (ORDER == 0) return bayer_interleave (R, G, B, A);
  if (ORDER == 1) return bayer_interleave (B, G, R, A);
  if (ORDER == 2) return bayer_interleave (A, R, G, B);
  return bayer_interleave (A, B, G, R);
}

static inline int bayer_order_of (int r, int g, int b) {
  if (r == 0 && g == 1 && b == 2) return 0;
  if (r == 2 && g == 1 && b == 0) return 1;
  if (r == 1 && g == 2 && b == 3) return 2;
  if (r == 3 && g == 2 && b == 1) return 3;
  return -1;
}

// ------------------------------------------------------------- fused epilogue
// Elements that follow bayer2rgb in a pipeline and are pure per-pixel functions fold
// into the demosaic kernel (5 B/px instead of 5+8+8, BASELINE.json config 5):
//   mode 1: per-byte-position LUT (burn, dodge, chromium, solarize, coloreffects xpro /
//           yellowblue, any composition of them): table word v = lut[0..3][v];
//   mode 2: a luma-mapped coloreffects preset (gstcoloreffects.c:331-344), with any
//           following per-channel LUTs already folded into the table by the host:
//           table word l = the finished output pixel for luma l.
struct BayerEpilogue {
  PackedLut table;
  uint32_t luma_weights;     // 54 / 183 / 19 at the R / G / B byte positions (dp4a operand)
  int mode;
};

#ifdef __CUDACC__
template <int MODE>
__device__ __forceinline__ uint4 bayer_epilogue (uint4 px, const uint32_t *tl, uint32_t wts) {
  if (MODE == 1) {
    px.x = lut_px (tl, px.x); px.y = lut_px (tl, px.y); px.z = lut_px (tl, px.z); px.w = lut_px (tl, px.w);
  } else if (MODE == 2) {
    // luma = (54 r + 183 g + 19 b) >> 8 in one dot product per pixel
    px.x = tl[(__dp4a (px.x, wts, 0u) >> 8) << 5];
    px.y = tl[(__dp4a (px.y, wts, 0u) >> 8) << 5];
    px.z = tl[(__dp4a (px.z, wts, 0u) >> 8) << 5];
    px.w = tl[(__dp4a (px.w, wts, 0u) >> 8) << 5];
  }
  return px;
}
#endif

// r/g/b offsets are the OUTPUT byte positions (before the rggb/gbrg red-blue swap of the
// demosaic: the swap only changes which computed plane lands where, the epilogue sees
// finished pixels).
static inline void bayer_build_epilogue (BayerEpilogue &e, int r_off, int g_off, int b_off,
    const uint8_t *luma_table768, const uint8_t (*lut)[256]) {
  e.mode = 0;
  e.luma_weights = 0;
  if (!luma_table768 && !lut) return;
  int a_off = 6 - r_off - g_off - b_off;
  if (luma_table768) {
    e.mode = 2;
    e.luma_weights = (54u << (8 * r_off)) | (183u << (8 * g_off)) | (19u << (8 * b_off));
    for (int l = 0; l < 256; l++) {
      uint32_t r = luma_table768[3 * l], g = luma_table768[3 * l + 1], b = luma_table768[3 * l + 2], a = 255;
      if (lut) { r = lut[r_off][r]; g = lut[g_off][g]; b = lut[b_off][b]; a = lut[a_off][255]; }
      e.table.w[l] = (r << (8 * r_off)) | (g << (8 * g_off)) | (b << (8 * b_off)) | (a << (8 * a_off));
    }
  } else {
    e.mode = 1;
    pack_lut4 (lut, e.table);
  }
}
