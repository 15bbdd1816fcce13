// pointops.cu - per-pixel point / LUT elements of gaudieffects for sm_100a.
//
// lut4: out.byte[c] = lut[c][in.byte[c]] over packed 4-byte pixels. Replaces
// gaudi_orc_burn (gst/gaudieffects/gstgaudieffectsorc.orc:1-25) and the static
// transform() loops of gstdodge.c:231-254, gstchromium.c:282-338 and
// gstsolarize.c:286-339, whose arithmetic moves into the host-side LUT builders
// below (libm cos for chromium stays on the host for bit-exactness).
//
// HBM-bound byte work (8 B/px): 128-bit coalesced streaming loads/stores, a
// persistent grid of CTAs (a multiple of the SM count) and a LANE-REPLICATED
// table in shared memory: entry v of lane l sits at word v*32+l, so a warp's 32
// random lookups always hit 32 different banks (no conflicts for random data),
// and one word carries the four channel tables, so a lookup is one LDS.32.
#include "lut.cuh"
#include "stream.cuh"
#include <math.h>
#include <string.h>

namespace {

constexpr int LUT_THREADS = 512;
constexpr int LUT_UNROLL = 4;            // 128-bit groups in flight per thread

__global__ void __launch_bounds__ (LUT_THREADS)
lut4_kernel (const uint4 *__restrict__ src, uint4 *__restrict__ dst, size_t n16,
    const uint32_t *__restrict__ src_tail, uint32_t *__restrict__ dst_tail, int ntail,
    const __grid_constant__ PackedLut lut)
{
  extern __shared__ uint32_t tab[];
  lut_fill (tab, lut);
  const uint32_t *tl = tab + (threadIdx.x & 31);
  const size_t stride = (size_t) gridDim.x * blockDim.x;
  size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + (LUT_UNROLL - 1) * stride < n16; i += LUT_UNROLL * stride) {
    uint4 v[LUT_UNROLL];
#pragma unroll
    for (int k = 0; k < LUT_UNROLL; k++) v[k] = ld_stream_v4 (src + i + k * stride);
#pragma unroll
    for (int k = 0; k < LUT_UNROLL; k++) {
      uint4 o;
      o.x = lut_px (tl, v[k].x); o.y = lut_px (tl, v[k].y);
      o.z = lut_px (tl, v[k].z); o.w = lut_px (tl, v[k].w);
      st_stream_v4 (dst + i + k * stride, o);
    }
  }
  for (; i < n16; i += stride) {
    uint4 v = ld_stream_v4 (src + i), o;
    o.x = lut_px (tl, v.x); o.y = lut_px (tl, v.y); o.z = lut_px (tl, v.z); o.w = lut_px (tl, v.w);
    st_stream_v4 (dst + i, o);
  }
  if (blockIdx.x == 0 && threadIdx.x < ntail) dst_tail[threadIdx.x] = lut_px (tl, src_tail[threadIdx.x]);
}

struct LutOp {                           // stream.cuh operator: the same table, the same lookup
  PackedLut lut;
  __device__ __forceinline__ void fill (uint32_t *tab) const { lut_fill (tab, lut); }
  __device__ __forceinline__ uint32_t apply (const uint32_t *tl, uint32_t px) const { return lut_px (tl, px); }
};

}  // namespace

int b200vf_persistent_grid (b200vf_ctx *ctx, int ctas_per_sm) { return ctx->sm_count * ctas_per_sm; }

B200VF_API int b200vf_lut4 (b200vf_ctx *ctx, const uint8_t *d_src, uint8_t *d_dst, size_t npix_total,
    const uint8_t lut[4][256], void *stream)
{
  B200VF_REQUIRE (ctx && d_src && d_dst && lut, B200VF_E_INVAL, "lut4: NULL argument");
  B200VF_REQUIRE (((uintptr_t) d_src) % 4 == 0 && ((uintptr_t) d_dst) % 4 == 0, B200VF_E_INVAL, "lut4: pixels must be 4-byte aligned");
  if (npix_total == 0) return B200VF_OK;
  PackedLut p;
  pack_lut4 (lut, p);
  // peel a head so that the body is 16-byte aligned on both sides when src/dst are congruent
  size_t head = 0;
  if (((uintptr_t) d_src) % 16 == ((uintptr_t) d_dst) % 16) head = ((16 - ((uintptr_t) d_src) % 16) % 16) / 4;
  if (head > npix_total) head = npix_total;
  bool aligned = (((uintptr_t) (d_src + 4 * head)) % 16 == 0) && (((uintptr_t) (d_dst + 4 * head)) % 16 == 0);
  cudaStream_t s = b200vf_stream (ctx, stream);
  const int smem = 256 * 32 * 4;
  if (int rc = b200vf_func_smem (ctx, (const void *) lut4_kernel, smem)) return rc;
  size_t body = aligned ? (npix_total - head) / 4 : 0;          // 128-bit groups
  size_t done = head + body * 4;
  // scalar leftovers: head pixels + tail pixels (<= 3 each), or everything when misaligned
  const uint32_t *s32 = reinterpret_cast<const uint32_t *> (d_src);
  uint32_t *d32 = reinterpret_cast<uint32_t *> (d_dst);
  if (!aligned) {
    // rare (buffers not congruent mod 16): run the vector loop over nothing and the scalar path over 32-bit words
    // by treating each pixel as a "tail"; do it in chunks of the CTA size through repeated launches
    size_t off = 0;
    while (off < npix_total) {
      int n = (int) ((npix_total - off) < (size_t) LUT_THREADS ? (npix_total - off) : LUT_THREADS);
      lut4_kernel<<<1, LUT_THREADS, smem, s>>> (nullptr, nullptr, 0, s32 + off, d32 + off, n, p);
      int rc = b200vf_launched (ctx, "lut4_unaligned");
      if (rc) return rc;
      off += n;
    }
    return B200VF_OK;
  }
  int ntail = (int) (npix_total - done);
  int rc;
  if (stream_enabled (ctx) && body >= 1024) {
    // TMA-fed ring (stream.cuh) over the 16-byte aligned body; the <= 3 tail pixels go with the head launch below
    LutOp op;
    op.lut = p;
    rc = stream_launch (ctx, d_src + 4 * head, d_dst + 4 * head, body * 16, op, s, "lut4_tma", 8);
    if (rc) return rc;
    if (ntail) {
      lut4_kernel<<<1, LUT_THREADS, smem, s>>> (nullptr, nullptr, 0, s32 + done, d32 + done, ntail, p);
      rc = b200vf_launched (ctx, "lut4_tail");
      if (rc) return rc;
    }
  } else {
    int grid = b200vf_persistent_grid (ctx, 2);
    size_t need = (body + LUT_THREADS - 1) / LUT_THREADS;
    if (need < (size_t) grid) grid = need ? (int) need : 1;
    // head pixels (if any) ride along as an extra tiny launch only when present
    lut4_kernel<<<grid, LUT_THREADS, smem, s>>> (reinterpret_cast<const uint4 *> (d_src + 4 * head),
        reinterpret_cast<uint4 *> (d_dst + 4 * head), body, s32 + done, d32 + done, ntail, p);
    rc = b200vf_launched (ctx, "lut4");
    if (rc) return rc;
  }
  if (head) {
    lut4_kernel<<<1, LUT_THREADS, smem, s>>> (nullptr, nullptr, 0, s32, d32, (int) head, p);
    rc = b200vf_launched (ctx, "lut4_head");
  }
  return rc;
}

// ------------------------------------------------------------ host LUT builders
// Each reproduces the reference element's per-byte arithmetic with the same C
// types; byte position 3 is what the element leaves in the x/alpha byte
// (burn transforms it, the others write 0). Pixels are little-endian u32:
// byte 0 = "blue" slot, 1 = "green", 2 = "red" of the reference's shifts.

B200VF_API int b200vf_lut_burn (int adjustment, uint8_t lut[4][256]) {
  B200VF_REQUIRE (lut && adjustment >= 0 && adjustment <= 256, B200VF_E_PROPERTY, "burn: adjustment %d not in [0,256]", adjustment);
  for (int c = 0; c < 256; c++) {
    // gaudi_orc_burn: addw, shruw 1, subb, shlw 7, divluw, subw (gstgaudieffectsorc.orc:12-23)
    uint16_t a = (uint16_t) ((uint16_t) (c + adjustment) >> 1);
    uint16_t t = (uint16_t) (((uint8_t) (255 - c)) << 7);
    unsigned div = a & 0xff, q;
    if (div == 0) q = 255;
    else { q = t / div; if (q > 255) q = 255; }
    uint8_t o = (uint8_t) (255 - q);
    lut[0][c] = lut[1][c] = lut[2][c] = lut[3][c] = o;     // x4: all four bytes
  }
  return B200VF_OK;
}

B200VF_API int b200vf_lut_dodge (uint8_t lut[4][256]) {
  B200VF_REQUIRE (lut, B200VF_E_INVAL, "dodge: NULL lut");
  for (int c = 0; c < 256; c++) {
    int v = (256 * c) / (256 - c);                          // gstdodge.c:243-245
    v = v > 255 ? 255 : (v < 0 ? 0 : v);
    lut[0][c] = lut[1][c] = lut[2][c] = (uint8_t) v;
    lut[3][c] = 0;                                          // :252 rebuilds the pixel without byte 3
  }
  return B200VF_OK;
}

B200VF_API int b200vf_lut_chromium (int edge_a, int edge_b, uint8_t lut[4][256]) {
  B200VF_REQUIRE (lut && edge_a >= 0 && edge_a <= 256 && edge_b >= 0 && edge_b <= 256, B200VF_E_PROPERTY,
      "chromium: edge-a %d / edge-b %d not in [0,256]", edge_a, edge_b);
  // setup_cos_table, gstchromium.c:282-291; built once, thread-safely (function-local static initialiser): the LUT
  // builders are called from any streaming thread
  struct CosTable {
    int v[1024];
    CosTable () {
      const float pi = 3.141582f;                             // sic, :102
      for (int angle = 0; angle < 1024; angle++) {
        float rad = ((float) angle / 512) * pi;
        v[angle] = (int) (cos (rad) * 512);
      }
    }
  };
  static const CosTable table;
  const int *cos_table = table.v;
  for (int c = 0; c < 256; c++) {
    int v = cos_table[((c + edge_a) + ((c * edge_b) / 2)) & 1023];   // :325-328
    if (v < 0) v = -v;
    v = v > 255 ? 255 : v;
    lut[0][c] = lut[1][c] = lut[2][c] = (uint8_t) v;
    lut[3][c] = 0;
  }
  return B200VF_OK;
}

B200VF_API int b200vf_lut_solarize (int threshold, int start, int end, uint8_t lut[4][256]) {
  B200VF_REQUIRE (lut && threshold >= 0 && threshold <= 256 && start >= 0 && start <= 256 && end >= 0 && end <= 256,
      B200VF_E_PROPERTY, "solarize: threshold/start/end %d/%d/%d not in [0,256]", threshold, start, end);
  // gstsolarize.c:286-339, type-for-type (gint / guint32 mix at :316-327)
  int period = 1, up_length = 1, down_length = 1;
  static const unsigned int ceiling = 255;
  if (end != start) period = end - start;
  if (threshold != start) up_length = threshold - start;
  if (threshold != end) down_length = end - threshold;
  for (int c = 0; c < 256; c++) {
    uint32_t color;
    int param = c;
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
    if (color > 255) color = 255;
    lut[0][c] = lut[1][c] = lut[2][c] = (uint8_t) color;
    lut[3][c] = 0;
  }
  return B200VF_OK;
}

B200VF_API int b200vf_lut_compose (const uint8_t first[4][256], const uint8_t second[4][256], uint8_t lut_out[4][256]) {
  B200VF_REQUIRE (first && second && lut_out, B200VF_E_INVAL, "lut_compose: NULL argument");
  uint8_t tmp[4][256];
  for (int c = 0; c < 4; c++)
    for (int v = 0; v < 256; v++) tmp[c][v] = second[c][first[c][v]];
  memcpy (lut_out, tmp, sizeof tmp);
  return B200VF_OK;
}
