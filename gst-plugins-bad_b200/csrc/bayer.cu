// bayer.cu - bayer2rgb / rgb2bayer for sm_100a.
//
// Replaces gst_bayer2rgb_process (gst/bayer/gstbayer2rgb.c:387-451) and its ORC
// programs; see include/b200vf.h for the contract and bayer.cuh for the algebra.
//
// Two kernels, one algebra:
//  * bayer2rgb_direct: any even width / any 4-byte-aligned pitch. One warp owns a
//    128-pixel-wide column strip of STRIP_ROWS rows; a lane owns 4 pixels, loads
//    one coalesced 32-bit word per input row, gets its left/right neighbour bytes
//    by warp shuffle, keeps the three upsampled rows of the 3x3 stencil in
//    registers while marching down (each input byte is fetched once per strip)
//    and emits one 128-bit store (4 RGBx pixels) per row.
//  * bayer2rgb_tma (bayer_tma.cu): the same math fed by TMA 2-D tiles staged in
//    shared memory; needs a 16-byte-aligned pitch.
// HBM-bound integer/byte work: no tensor cores.
#include "bayer.cuh"
#include "lut.cuh"

namespace {

constexpr int STRIP_ROWS = 16;      // rows one warp marches down (2 halo rows re-read per strip)
constexpr int WARPS_PER_CTA = 8;
constexpr int PREFETCH = 4;         // input rows in flight per lane

struct BayerParams {
  const uint8_t *src;   // virtual pointer to GLOBAL row 0 of frame 0 (only touched rows must exist)
  uint8_t *dst;         // pointer to the first OUTPUT row of this call (global row `row0`), frame 0
  size_t src_frame_stride, dst_frame_stride;
  int src_stride, dst_stride;
  int width, full_height;
  int row0, rows;       // output rows [row0, row0+rows)
  int first_is_gr;      // grbg / gbrg start with the "gr" merge (gstbayer2rgb.c:422-427)
};

struct RawRow { uint32_t s, e; };

__device__ __forceinline__ RawRow bayer_issue (const uint8_t *rp, int x0, int width, int lane, bool active) {
  RawRow r;
  r.s = active ? ldg_u32 (rp + x0) : 0u;
  r.e = 0u;
  // the warp's outer neighbour bytes: lane 0 fetches the word to its left,
  // lane 31 the word to its right (L1/L2 hits: the neighbouring warp streams them)
  if (lane == 0 && x0 > 0 && active) r.e = ldg_u32 (rp + x0 - 4);
  if (lane == 31 && x0 + 4 < width) r.e = ldg_u32 (rp + x0 + 4);
  return r;
}

__device__ __forceinline__ BayerRow bayer_finish (RawRow r, int lane, uint32_t selL, uint32_t selR) {
  uint32_t prev = __shfl_up_sync (0xffffffffu, r.s, 1);
  uint32_t next = __shfl_down_sync (0xffffffffu, r.s, 1);
  if (lane == 0) prev = r.e;
  if (lane == 31) next = r.e;
  return bayer_upsample (prev, r.s, next, selL, selR);
}

// global row feeding the stencil below output row j (gstbayer2rgb.c:429-448):
// row h is never upsampled, its ring slot still holds row h-4 (row 1 when h == 3)
__device__ __forceinline__ int bayer_down_row (int j, int h) {
  return (j + 1 < h) ? j + 1 : (h >= 4 ? h - 4 : 1);
}

template <int ORDER, bool VEC16, int MODE>
__global__ void __launch_bounds__ (32 * WARPS_PER_CTA)
bayer2rgb_direct_kernel (const BayerParams p, const __grid_constant__ BayerEpilogue epi)
{
  extern __shared__ uint32_t epi_tab[];
  if (MODE != 0) lut_fill (epi_tab, epi.table);
  const uint32_t *tl = epi_tab + (threadIdx.x & 31);
  const int lane = threadIdx.x;
  const int x0 = (blockIdx.x * 32 + lane) * 4;
  const int strip = blockIdx.y * WARPS_PER_CTA + threadIdx.y;
  const int j0 = p.row0 + strip * STRIP_ROWS;
  const int jend = min (j0 + STRIP_ROWS, p.row0 + p.rows);
  if (j0 >= jend) return;                                   // warp-uniform
  const int h = p.full_height, w = p.width;
  const bool active = x0 < w;
  const int v = min (4, w - x0);
  const uint32_t selL = bayer_selL (x0);
  const uint32_t selR = bayer_selR (active && x0 + 4 >= w, v);
  const uint8_t *src = p.src + (size_t) blockIdx.z * p.src_frame_stride;
  uint8_t *dst = p.dst + (size_t) blockIdx.z * p.dst_frame_stride;

  // prime the stencil: rows j0-1 (row 1 mirrors the top edge, :432-433) and j0
  RawRow ru = bayer_issue (src + (size_t) (j0 == 0 ? 1 : j0 - 1) * p.src_stride, x0, w, lane, active);
  RawRow rc = bayer_issue (src + (size_t) j0 * p.src_stride, x0, w, lane, active);
  BayerRow u = bayer_finish (ru, lane, selL, selR);
  BayerRow c = bayer_finish (rc, lane, selL, selR);

  for (int j = j0; j < jend; j += PREFETCH) {
    RawRow raw[PREFETCH];
#pragma unroll
    for (int k = 0; k < PREFETCH; k++) {
      int jj = min (j + k, jend - 1);
      raw[k] = bayer_issue (src + (size_t) bayer_down_row (jj, h) * p.src_stride, x0, w, lane, active);
    }
#pragma unroll
    for (int k = 0; k < PREFETCH; k++) {
      const int jj = j + k;
      if (jj < jend) {                                      // warp-uniform
        BayerRow d = bayer_finish (raw[k], lane, selL, selR);
        uint32_t R, G, B;
        bayer_merge (u, c, d, ((jj & 1) != 0) != (p.first_is_gr != 0), R, G, B);
        uint4 px = bayer_pack<ORDER> (R, G, B, 0xffffffffu);
        px = bayer_epilogue<MODE> (px, tl, epi.luma_weights);
        uint8_t *o = dst + (size_t) (jj - p.row0) * p.dst_stride + (size_t) x0 * 4;
        if (active) {
          if (VEC16 && v == 4) st_stream_v4 (o, px);
          else {
            st_stream_v2 (o, make_uint2 (px.x, px.y));
            if (v == 4) st_stream_v2 (o + 8, make_uint2 (px.z, px.w));
          }
        }
        u = c;
        c = d;
      }
    }
  }
}

template <int ORDER, int MODE>
int launch_direct_mode (b200vf_ctx *ctx, const BayerParams &p, const BayerEpilogue &epi, int nframes, cudaStream_t s) {
  dim3 block (32, WARPS_PER_CTA);
  int strips = (p.rows + STRIP_ROWS - 1) / STRIP_ROWS;
  dim3 grid ((p.width + 127) / 128, (strips + WARPS_PER_CTA - 1) / WARPS_PER_CTA, nframes);
  bool vec16 = (((uintptr_t) p.dst) % 16 == 0) && (p.dst_stride % 16 == 0) && (p.dst_frame_stride % 16 == 0);
  const int smem = MODE ? LUT_SMEM_BYTES : 0;
  if (vec16) bayer2rgb_direct_kernel<ORDER, true, MODE><<<grid, block, smem, s>>> (p, epi);
  else bayer2rgb_direct_kernel<ORDER, false, MODE><<<grid, block, smem, s>>> (p, epi);
  return b200vf_launched (ctx, MODE ? "bayer2rgb_direct_fused" : "bayer2rgb_direct");
}
template <int ORDER>
int launch_direct (b200vf_ctx *ctx, const BayerParams &p, const BayerEpilogue &epi, int nframes, cudaStream_t s) {
  switch (epi.mode) {
    case 0: return launch_direct_mode<ORDER, 0> (ctx, p, epi, nframes, s);
    case 1: return launch_direct_mode<ORDER, 1> (ctx, p, epi, nframes, s);
    default: return launch_direct_mode<ORDER, 2> (ctx, p, epi, nframes, s);
  }
}

template <int ORDER>
int set_smem_order (b200vf_ctx *ctx) {
  int rc;
  if ((rc = b200vf_func_smem (ctx, (const void *) bayer2rgb_direct_kernel<ORDER, true, 1>, LUT_SMEM_BYTES))) return rc;
  if ((rc = b200vf_func_smem (ctx, (const void *) bayer2rgb_direct_kernel<ORDER, false, 1>, LUT_SMEM_BYTES))) return rc;
  if ((rc = b200vf_func_smem (ctx, (const void *) bayer2rgb_direct_kernel<ORDER, true, 2>, LUT_SMEM_BYTES))) return rc;
  return b200vf_func_smem (ctx, (const void *) bayer2rgb_direct_kernel<ORDER, false, 2>, LUT_SMEM_BYTES);
}
int bayer_direct_set_smem (b200vf_ctx *ctx) {
  int rc;
  if ((rc = set_smem_order<0> (ctx)) || (rc = set_smem_order<1> (ctx)) || (rc = set_smem_order<2> (ctx))) return rc;
  return set_smem_order<3> (ctx);
}

}  // namespace

int b200vf_bayer2rgb_tma_launch (b200vf_ctx *ctx, const uint8_t *d_src, int src_stride, size_t src_frame_stride,
    uint8_t *d_dst, int dst_stride, size_t dst_frame_stride, int width, int full_height, int row0, int rows, int nframes,
    int order, int first_is_gr, const BayerEpilogue &epi, cudaStream_t s);
bool b200vf_bayer2rgb_tma_usable (b200vf_ctx *ctx, const uint8_t *d_src, int src_stride, size_t src_frame_stride,
    const uint8_t *d_dst, int dst_stride, size_t dst_frame_stride, int width, int full_height, int row0, int rows);

static int bayer_common (b200vf_ctx *ctx, const uint8_t *d_src, int src_stride, size_t src_frame_stride,
    uint8_t *d_dst, int dst_stride, size_t dst_frame_stride, int width, int full_height, int row0, int rows,
    int nframes, int pattern, int r_off, int g_off, int b_off, bool allow_tma, const uint8_t *luma_table768,
    const uint8_t (*lut)[256], void *stream)
{
  B200VF_REQUIRE (ctx && d_src && d_dst, B200VF_E_INVAL, "bayer2rgb: NULL argument");
  B200VF_REQUIRE (width >= 4 && (width & 1) == 0 && full_height >= 3, B200VF_E_INVAL,
      "bayer2rgb: %dx%d is outside the reference's domain (even width >= 4, height >= 3)", width, full_height);
  B200VF_REQUIRE (pattern >= 0 && pattern <= 3, B200VF_E_INVAL, "bayer2rgb: pattern %d", pattern);
  B200VF_REQUIRE (nframes >= 1 && rows >= 1 && row0 >= 0 && row0 + rows <= full_height, B200VF_E_INVAL,
      "bayer2rgb: rows [%d,%d) of %d, %d frames", row0, row0 + rows, full_height, nframes);
  B200VF_REQUIRE (src_stride >= width && src_stride % 4 == 0 && ((uintptr_t) d_src) % 4 == 0 &&
      src_frame_stride % 4 == 0, B200VF_E_INVAL, "bayer2rgb: source pitch %d / base must be 4-byte aligned", src_stride);
  B200VF_REQUIRE (dst_stride >= 4 * width && dst_stride % 8 == 0 && ((uintptr_t) d_dst) % 8 == 0 &&
      dst_frame_stride % 8 == 0, B200VF_E_INVAL, "bayer2rgb: destination pitch %d / base must be 8-byte aligned", dst_stride);
  const int out_r = r_off, out_g = g_off, out_b = b_off;   // where R,G,B really land in the output pixel
  // RGGB and GBRG swap red and blue, GRBG and GBRG start with the gr row (gstbayer2rgb.c:399-427)
  if (pattern == 3 || pattern == 1) { int t = r_off; r_off = b_off; b_off = t; }
  int first_is_gr = (pattern == 2 || pattern == 1);
  int order = bayer_order_of (r_off, g_off, b_off);
  B200VF_REQUIRE (order >= 0, B200VF_E_UNSUPPORTED,
      "bayer2rgb: offsets (%d,%d,%d) are none of the four layouts the reference dispatches on", r_off, g_off, b_off);
  cudaStream_t s = b200vf_stream (ctx, stream);
  BayerEpilogue epi;              // 1 KB table: built per call, passed by value to the kernel
  bayer_build_epilogue (epi, out_r, out_g, out_b, luma_table768, lut);
  if (epi.mode) {
    int rc = bayer_direct_set_smem (ctx);
    if (rc) return rc;
  }

  if (allow_tma && ctx->variant != 1 &&
      b200vf_bayer2rgb_tma_usable (ctx, d_src, src_stride, src_frame_stride, d_dst, dst_stride, dst_frame_stride,
          width, full_height, row0, rows))
    return b200vf_bayer2rgb_tma_launch (ctx, d_src, src_stride, src_frame_stride, d_dst, dst_stride,
        dst_frame_stride, width, full_height, row0, rows, nframes, order, first_is_gr, epi, s);
  B200VF_REQUIRE (ctx->variant != 2, B200VF_E_UNSUPPORTED,
      "bayer2rgb: the TMA variant was forced but this geometry needs the direct kernel");

  BayerParams p;
  p.src = d_src - (size_t) row0 * src_stride;   // virtual global row 0
  p.dst = d_dst;
  p.src_frame_stride = src_frame_stride;
  p.dst_frame_stride = dst_frame_stride;
  p.src_stride = src_stride;
  p.dst_stride = dst_stride;
  p.width = width;
  p.full_height = full_height;
  p.row0 = row0;
  p.rows = rows;
  p.first_is_gr = first_is_gr;
  switch (order) {
    case 0: return launch_direct<0> (ctx, p, epi, nframes, s);
    case 1: return launch_direct<1> (ctx, p, epi, nframes, s);
    case 2: return launch_direct<2> (ctx, p, epi, nframes, s);
    default: return launch_direct<3> (ctx, p, epi, nframes, s);
  }
}

B200VF_API int b200vf_bayer2rgb (b200vf_ctx *ctx, const uint8_t *d_src, int src_stride, size_t src_frame_stride,
    uint8_t *d_dst, int dst_stride, size_t dst_frame_stride, int width, int height, int nframes,
    int pattern, int r_off, int g_off, int b_off, void *stream)
{
  return bayer_common (ctx, d_src, src_stride, src_frame_stride, d_dst, dst_stride, dst_frame_stride,
      width, height, 0, height, nframes, pattern, r_off, g_off, b_off, true, nullptr, nullptr, stream);
}

B200VF_API int b200vf_bayer2rgb_fused (b200vf_ctx *ctx, const uint8_t *d_src, int src_stride, size_t src_frame_stride,
    uint8_t *d_dst, int dst_stride, size_t dst_frame_stride, int width, int height, int nframes,
    int pattern, int r_off, int g_off, int b_off, const uint8_t *luma_table768,
    const uint8_t lut[4][256], void *stream)
{
  return bayer_common (ctx, d_src, src_stride, src_frame_stride, d_dst, dst_stride, dst_frame_stride,
      width, height, 0, height, nframes, pattern, r_off, g_off, b_off, true, luma_table768, lut, stream);
}

B200VF_API int b200vf_bayer2rgb_shard (b200vf_ctx *ctx, const uint8_t *d_src, int src_stride, size_t src_frame_stride,
    uint8_t *d_dst, int dst_stride, size_t dst_frame_stride, int width, int full_height, int row0, int rows,
    int nframes, int pattern, int r_off, int g_off, int b_off, void *stream)
{
  // the bottom rule needs global row full_height-4 (row 1 when full_height == 3) next to the last row
  if (row0 + rows == full_height && row0 > 0)
    B200VF_REQUIRE (rows >= 4, B200VF_E_INVAL, "bayer2rgb_shard: the last shard needs >= 4 rows (has %d)", rows);
  if (row0 == 0 && rows < full_height)
    B200VF_REQUIRE (rows >= 2, B200VF_E_INVAL, "bayer2rgb_shard: the first shard needs >= 2 rows (has %d)", rows);
  return bayer_common (ctx, d_src, src_stride, src_frame_stride, d_dst, dst_stride, dst_frame_stride,
      width, full_height, row0, rows, nframes, pattern, r_off, g_off, b_off, true, nullptr, nullptr, stream);
}

B200VF_API int b200vf_bayer2rgb_shard_fused (b200vf_ctx *ctx, const uint8_t *d_src, int src_stride, size_t src_frame_stride,
    uint8_t *d_dst, int dst_stride, size_t dst_frame_stride, int width, int full_height, int row0, int rows,
    int nframes, int pattern, int r_off, int g_off, int b_off, const uint8_t *luma_table768,
    const uint8_t lut[4][256], void *stream)
{
  if (row0 + rows == full_height && row0 > 0)
    B200VF_REQUIRE (rows >= 4, B200VF_E_INVAL, "bayer2rgb_shard: the last shard needs >= 4 rows (has %d)", rows);
  if (row0 == 0 && rows < full_height)
    B200VF_REQUIRE (rows >= 2, B200VF_E_INVAL, "bayer2rgb_shard: the first shard needs >= 2 rows (has %d)", rows);
  return bayer_common (ctx, d_src, src_stride, src_frame_stride, d_dst, dst_stride, dst_frame_stride,
      width, full_height, row0, rows, nframes, pattern, r_off, g_off, b_off, true, luma_table768, lut, stream);
}

// ------------------------------------------------------------------ rgb2bayer
// gst/bayer/gstrgb2bayer.c:254-267: dest[i] = byte 3 / 1 / 2 of the ARGB pixel
// depending on (row, column) parity vs the pattern. HBM-bound (4 B read, 1 B written per
// pixel): a lane owns 16 pixels of a row - four independent 128-bit streaming loads in
// flight, one PRMT per pixel pair, one 128-bit store.
namespace {
__device__ __forceinline__ uint32_t rgb2bayer_pick4 (uint4 v, uint32_t sel_pair) {
  // bytes (even pixel, odd pixel) of two pixel pairs -> 4 mosaic bytes
  return PRMT (PRMT (v.x, v.y, sel_pair), PRMT (v.z, v.w, sel_pair), 0x5410);
}
__global__ void __launch_bounds__ (256)
rgb2bayer_kernel (const uint8_t *src, int src_stride, size_t src_fs, uint8_t *dst, int dst_stride, size_t dst_fs,
    int width, int height, int pattern)
{
  const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 16;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  if (x0 >= width || j >= height) return;
  const uint8_t *sp = src + (size_t) blockIdx.z * src_fs + (size_t) j * src_stride + (size_t) x0 * 4;
  uint8_t *dp = dst + (size_t) blockIdx.z * dst_fs + (size_t) j * dst_stride + x0;
  // byte picked at even / odd columns of this row
  int sel[2];
#pragma unroll
  for (int par = 0; par < 2; par++) {
    int is_blue = ((j & 1) << 1) | par;
    sel[par] = (is_blue == pattern) ? 3 : (((is_blue ^ 3) == pattern) ? 1 : 2);
  }
  const int n = min (16, width - x0);
  if (n == 16 && (((uintptr_t) sp) & 15) == 0 && (((uintptr_t) dp) & 15) == 0) {
    const uint32_t sel_pair = (uint32_t) sel[0] | ((uint32_t) (4 + sel[1]) << 4);     // byte 0 <- even pixel, byte 1 <- odd pixel
    const uint4 v0 = ld_stream_v4 (sp), v1 = ld_stream_v4 (sp + 16), v2 = ld_stream_v4 (sp + 32), v3 = ld_stream_v4 (sp + 48);
    st_stream_v4 (dp, make_uint4 (rgb2bayer_pick4 (v0, sel_pair), rgb2bayer_pick4 (v1, sel_pair),
        rgb2bayer_pick4 (v2, sel_pair), rgb2bayer_pick4 (v3, sel_pair)));
  } else {
    for (int i = 0; i < n; i++) dp[i] = sp[4 * i + sel[i & 1]];
  }
}
}  // namespace

B200VF_API int b200vf_rgb2bayer (b200vf_ctx *ctx, const uint8_t *d_src, int src_stride, size_t src_frame_stride,
    uint8_t *d_dst, int dst_stride, size_t dst_frame_stride, int width, int height, int nframes,
    int pattern, void *stream)
{
  B200VF_REQUIRE (ctx && d_src && d_dst && width > 0 && height > 0 && nframes > 0, B200VF_E_INVAL, "rgb2bayer: bad argument");
  B200VF_REQUIRE (pattern >= 0 && pattern <= 3, B200VF_E_INVAL, "rgb2bayer: pattern %d", pattern);
  B200VF_REQUIRE (src_stride >= 4 * width && dst_stride >= width, B200VF_E_INVAL, "rgb2bayer: strides");
  B200VF_REQUIRE (dst_stride % 4 == 0 && ((uintptr_t) d_dst) % 4 == 0 && dst_frame_stride % 4 == 0, B200VF_E_INVAL,
      "rgb2bayer: destination pitch/base must be 4-byte aligned");
  dim3 block (64, 4);                                      // 1024 pixels x 4 rows per CTA
  dim3 grid ((width + 1023) / 1024, (height + 3) / 4, nframes);
  rgb2bayer_kernel<<<grid, block, 0, b200vf_stream (ctx, stream)>>> (d_src, src_stride, src_frame_stride,
      d_dst, dst_stride, dst_frame_stride, width, height, pattern);
  return b200vf_launched (ctx, "rgb2bayer");
}
