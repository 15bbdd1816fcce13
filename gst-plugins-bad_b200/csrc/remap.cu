// remap.cu - geometrictransform gather for sm_100a.
//
// Replaces the per-pixel loop of gst_geometric_transform_transform_frame
// (gst/geometrictransform/gstgeometrictransform.c:226-293) + do_map (:167-207):
// the frame is cleared to the fill value and every output pixel copies
// pixel_stride bytes from the input pixel its map entry truncates to. The double map
// and the off-edge policy are resolved on the host into an int32 table
// (gt_maps.cpp); here: coalesced 128-bit index loads, 4-byte read-only gathers
// (the zoomed source region of smooth maps stays in the 126 MB L2), 128-bit streaming
// stores, and the clear folded into the same pass (no separate memset).
#include "common.cuh"
#include <stdlib.h>
#include <string.h>

namespace {

__device__ __forceinline__ int4 ld_idx4 (const int32_t *p) {
  int4 r;
  asm volatile ("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
      : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

// 4-byte pixels, rows contiguous (row_stride == 4*width). A warp takes 128 consecutive output pixels per
// iteration, lane l the pixels l, 32+l, 64+l, 96+l: every gather instruction then reads the source
// positions of 32 NEIGHBOURING output pixels, i.e. a dense run of the source row (4-5 sectors, all bytes
// used). Giving each lane 4 consecutive pixels instead (128-bit index loads and stores) makes each gather
// a stride-4 walk over the same sectors four times: 14 sectors per request in the ncu capture of that
// version (profiles/), and the kernel was bound by L1/L2 transactions, not by HBM.
__global__ void __launch_bounds__ (256)
remap4_kernel (const uint32_t *__restrict__ src, uint32_t *__restrict__ dst, const int32_t *__restrict__ index,
    size_t npix, size_t frame_px, uint32_t fill)
{
  const uint32_t *s = src + (size_t) blockIdx.y * frame_px;
  uint32_t *d = dst + (size_t) blockIdx.y * frame_px;
  const int lane = threadIdx.x & 31;
  const size_t warps = (size_t) gridDim.x * (blockDim.x / 32);
  for (size_t base = ((size_t) blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5)) * 128; base < npix; base += warps * 128) {
    int ix[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const size_t i = base + j * 32 + lane;
      ix[j] = i < npix ? (int) ldg_u32 (index + i) : -2;
    }
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; j++) o[j] = ix[j] >= 0 ? __ldg (s + ix[j]) : fill;
#pragma unroll
    for (int j = 0; j < 4; j++) if (ix[j] != -2) st_stream_u32 (d + base + j * 32 + lane, o[j]);
  }
}

// any pixel stride (1,2,3,4) and padded rows: one output pixel per thread
__global__ void __launch_bounds__ (256)
remap_generic_kernel (const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, const int32_t *__restrict__ index,
    int width, int height, int ps, int row_stride, size_t frame_stride, uint32_t fill)
{
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  const uint8_t *s = src + (size_t) blockIdx.z * frame_stride;
  uint8_t *d = dst + (size_t) blockIdx.z * frame_stride + (size_t) y * row_stride;
  if (x < width) {
    int ix = index[(size_t) y * width + x];
    uint8_t *o = d + (size_t) x * ps;
    if (ix >= 0) {
      const uint8_t *in = s + (size_t) (ix / width) * row_stride + (size_t) (ix % width) * ps;
      for (int b = 0; b < ps; b++) o[b] = in[b];
    } else {
      // the cleared frame: memset 0, or the AYUV pattern written as 32-bit words from the frame start (:244-252)
      for (int b = 0; b < ps; b++) o[b] = (uint8_t) (fill >> (8 * ((x * ps + b) & 3)));
    }
  }
  // row padding belongs to the cleared frame too (memset covers map[0].size)
  const int pad0 = width * ps;
  for (int b = pad0 + x; b < row_stride; b += gridDim.x * blockDim.x) d[b] = (uint8_t) (fill >> (8 * (b & 3)));
}

// ---- packed index table -----------------------------------------------------------------
// The int32 table costs as much HBM traffic as the pixels it moves (4 B/px next to 4 B read + 4 B written).
// Smooth maps (everything but diffuse) move the source position by a few pixels from one output pixel to
// the next, so the table is stored per chunk of 8 consecutive pixels of one output row as the source
// position (tx, ty) of the first pixel (2 x int16 in one word) plus one byte per pixel holding the step from
// the previous pixel, (dtx + 8) | (dty + 8) << 4 with both steps in [-8, 7]: 12 bytes per 8 pixels = 1.5 B/px.
// Chunks that cannot be coded (an ignored pixel, a discontinuity of the map) keep their 8 int32 entries in a
// side array (base word = ~slot < 0). One thread decodes one chunk: the 4+4 nibble prefix sums are two
// multiplies by 0x01010101 (bytes never exceed 8 * 15), no cross-lane traffic.
constexpr int CHUNK = 8;

__device__ __forceinline__ int byte_of (uint32_t v, int k) { return (int) ((v >> (8 * k)) & 0xffu); }

// ALIGNED: width % 8 == 0, so chunk c covers output pixels 8c .. 8c+7 and the 256 entries a warp decodes per
// iteration are 256 consecutive output pixels (no per-chunk offset table, no ragged row ends).
template <bool ALIGNED>
__global__ void __launch_bounds__ (256, 4)
remap4_packed_kernel (const uint32_t *__restrict__ src, uint32_t *__restrict__ dst, const int32_t *__restrict__ bases,
    const uint2 *__restrict__ steps, const int32_t *__restrict__ raw, int w, int chunks_per_row, uint32_t n_chunks,
    size_t frame_px, uint32_t fill)
{
  __shared__ int s_ix[8][32 * CHUNK];                                     // per warp: the 256 decoded source positions
  __shared__ int s_off[ALIGNED ? 1 : 8][32];                              //           and each chunk's first output pixel
  const uint32_t *s = src + (size_t) blockIdx.y * frame_px;
  uint32_t *d = dst + (size_t) blockIdx.y * frame_px;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const uint32_t stride = gridDim.x * (blockDim.x / 32) * 32;            // chunk numbers fit 32 bits (w, h <= 32767): no 64-bit division
  uint32_t c = (blockIdx.x * (blockDim.x / 32) + wib) * 32 + lane;        // this lane decodes chunk c (8 pixels) ...
  const int w1 = w + 1;
  // table words of the next iteration are fetched while this one gathers (one dependent memory round trip
  // per iteration instead of two)
  int32_t base = c < n_chunks ? __ldg (bases + c) : -1;
  uint2 st = c < n_chunks ? __ldg (steps + c) : make_uint2 (0u, 0u);
  for (; c - lane < n_chunks; c += stride) {
    const uint32_t cn = c + stride;
    const int32_t base_n = cn < n_chunks ? __ldg (bases + cn) : -1;
    const uint2 st_n = cn < n_chunks ? __ldg (steps + cn) : make_uint2 (0u, 0u);
    int ix[CHUNK];
    int off = 0;
    if (c < n_chunks) {
      if (base < 0) {                                                     // raw chunk
        const int4 a = ld_idx4 (raw + (size_t) (~base) * CHUNK), b = ld_idx4 (raw + (size_t) (~base) * CHUNK + 4);
        ix[0] = a.x; ix[1] = a.y; ix[2] = a.z; ix[3] = a.w; ix[4] = b.x; ix[5] = b.y; ix[6] = b.z; ix[7] = b.w;
      } else {
        // byte k of px/py (qx/qy) = sum of the biased steps 0..k (4..4+k) of this chunk
        const uint32_t px = (st.x & 0x0f0f0f0fu) * 0x01010101u, py = ((st.x >> 4) & 0x0f0f0f0fu) * 0x01010101u;
        const uint32_t qx = (st.y & 0x0f0f0f0fu) * 0x01010101u + (px >> 24) * 0x01010101u;
        const uint32_t qy = ((st.y >> 4) & 0x0f0f0f0fu) * 0x01010101u + (py >> 24) * 0x01010101u;
        const int b0 = (base >> 16) * w + (base & 0xffff) - 8 * w1;       // position of pixel 0 minus one step bias
#pragma unroll
        for (int k = 0; k < 4; k++) {
          ix[k] = b0 - 8 * k * w1 + byte_of (py, k) * w + byte_of (px, k);
          ix[4 + k] = b0 - 8 * (4 + k) * w1 + byte_of (qy, k) * w + byte_of (qx, k);
        }
      }
      if (!ALIGNED) {
        const int y = (int) (c / (uint32_t) chunks_per_row), x0 = (int) (c - (uint32_t) y * chunks_per_row) * CHUNK;
        off = y * w + x0;
#pragma unroll
        for (int k = 0; k < CHUNK; k++) if (x0 + k >= w) ix[k] = -2;      // ragged row end: no such pixel
      }
    } else {
#pragma unroll
      for (int k = 0; k < CHUNK; k++) ix[k] = -2;
    }
    // ... and the warp gathers them pixel-interleaved: lane l takes entries l, 32+l, ... so that one gather
    // instruction reads the source of 32 neighbouring output pixels (see remap4_kernel)
    __syncwarp ();
    *reinterpret_cast<int4 *> (&s_ix[wib][lane * CHUNK]) = make_int4 (ix[0], ix[1], ix[2], ix[3]);
    *reinterpret_cast<int4 *> (&s_ix[wib][lane * CHUNK + 4]) = make_int4 (ix[4], ix[5], ix[6], ix[7]);
    if (!ALIGNED) s_off[wib][lane] = off;
    __syncwarp ();
    int jx[CHUNK];
    uint32_t o[CHUNK];
#pragma unroll
    for (int j = 0; j < CHUNK; j++) jx[j] = s_ix[wib][j * 32 + lane];
#pragma unroll
    for (int j = 0; j < CHUNK; j++) o[j] = jx[j] >= 0 ? __ldg (s + jx[j]) : fill;
    uint32_t *out = d + (size_t) (c - lane) * CHUNK + lane;               // ALIGNED: entry p of the warp = output pixel 8 * c0 + p
#pragma unroll
    for (int j = 0; j < CHUNK; j++) {
      const int p = j * 32 + lane;
      if (jx[j] == -2) continue;
      if (ALIGNED) st_stream_u32 (out + j * 32, o[j]);
      else st_stream_u32 (d + s_off[wib][p >> 3] + (p & 7), o[j]);
    }
    base = base_n; st = st_n;
  }
}

struct PackedLayout { size_t bases, steps, raw, n_chunks; int chunks_per_row; };
PackedLayout packed_layout (int w, int h) {
  PackedLayout L;
  L.chunks_per_row = (w + CHUNK - 1) / CHUNK;
  L.n_chunks = (size_t) L.chunks_per_row * h;
  L.bases = 64;                                                           // after a 64-byte file header
  L.steps = L.bases + ((L.n_chunks * 4 + 63) & ~(size_t) 63);
  L.raw = L.steps + ((L.n_chunks * 8 + 63) & ~(size_t) 63);
  return L;
}
struct PackedFileHeader { uint32_t magic, version; int32_t w, h, chunks_per_row, n_raw; uint64_t total; };
const uint32_t PACKED_MAGIC = 0x50523242u;                                // "B2RP"

}  // namespace

B200VF_API size_t b200vf_gt_packed_bound (int width, int height) {
  if (width <= 0 || height <= 0) return 0;
  PackedLayout L = packed_layout (width, height);
  return L.raw + L.n_chunks * CHUNK * 4;
}

B200VF_API int b200vf_gt_pack_index (const int32_t *index, int width, int height, void *packed, size_t capacity,
    size_t *used, size_t *raw_chunks)
{
  B200VF_REQUIRE (index && packed && width > 0 && height > 0, B200VF_E_INVAL, "gt_pack_index: bad argument");
  B200VF_REQUIRE (width <= 32767 && height <= 32767, B200VF_E_UNSUPPORTED, "gt_pack_index: %dx%d exceeds 16-bit positions", width, height);
  const PackedLayout L = packed_layout (width, height);
  B200VF_REQUIRE (capacity >= L.raw, B200VF_E_INVAL, "gt_pack_index: capacity %zu < %zu", capacity, L.raw);
  uint8_t *base = static_cast<uint8_t *> (packed);
  int32_t *bases = reinterpret_cast<int32_t *> (base + L.bases);
  uint8_t *steps = base + L.steps;
  int32_t *raw = reinterpret_cast<int32_t *> (base + L.raw);
  memset (base, 0, L.raw);
  size_t n_raw = 0;
  for (int y = 0; y < height; y++) {
    for (int ci = 0; ci < L.chunks_per_row; ci++) {
      const size_t c = (size_t) y * L.chunks_per_row + ci;
      const int x0 = ci * CHUNK, n = (width - x0 < CHUNK) ? width - x0 : CHUNK;
      const int32_t *ix = index + (size_t) y * width + x0;
      uint8_t *st = steps + c * CHUNK;
      bool ok = ix[0] >= 0;
      int px = 0, py = 0;
      if (ok) { px = ix[0] % width; py = ix[0] / width; }
      for (int i = 1; ok && i < n; i++) {
        if (ix[i] < 0) { ok = false; break; }
        const int tx = ix[i] % width, ty = ix[i] / width;
        const int dx = tx - px, dy = ty - py;
        if (dx < -8 || dx > 7 || dy < -8 || dy > 7) { ok = false; break; }
        st[i] = (uint8_t) ((dx + 8) | ((dy + 8) << 4));
        px = tx; py = ty;
      }
      if (ok) {
        st[0] = 0x88;
        for (int i = n; i < CHUNK; i++) st[i] = 0x88;
        bases[c] = (ix[0] % width) | ((ix[0] / width) << 16);
      } else {
        if (L.raw + (n_raw + 1) * CHUNK * 4 > capacity) {
          b200vf_set_error ("gt_pack_index: capacity %zu too small (b200vf_gt_packed_bound)", capacity);
          return B200VF_E_INVAL;
        }
        memset (st, 0x88, CHUNK);
        int32_t *r = raw + n_raw * CHUNK;
        for (int i = 0; i < CHUNK; i++) r[i] = i < n ? ix[i] : -1;
        bases[c] = ~(int32_t) n_raw;
        n_raw++;
      }
    }
  }
  PackedFileHeader fh;
  fh.magic = PACKED_MAGIC; fh.version = 2; fh.w = width; fh.h = height; fh.chunks_per_row = L.chunks_per_row;
  fh.n_raw = (int32_t) n_raw; fh.total = L.raw + n_raw * CHUNK * 4;
  memcpy (base, &fh, sizeof fh);
  if (used) *used = (size_t) fh.total;
  if (raw_chunks) *raw_chunks = n_raw;
  return B200VF_OK;
}

// host decoder of the packed table (tests, tooling): the inverse of b200vf_gt_pack_index
B200VF_API int b200vf_gt_unpack_index (const void *packed, size_t size, int width, int height, int32_t *index) {
  B200VF_REQUIRE (packed && index && width > 0 && height > 0 && size >= 64, B200VF_E_INVAL, "gt_unpack_index: bad argument");
  PackedFileHeader fh;
  memcpy (&fh, packed, sizeof fh);
  B200VF_REQUIRE (fh.magic == PACKED_MAGIC && fh.version == 2 && fh.w == width && fh.h == height && fh.total <= size,
      B200VF_E_INVAL, "gt_unpack_index: not a packed table for %dx%d", width, height);
  const PackedLayout L = packed_layout (width, height);
  const uint8_t *base = static_cast<const uint8_t *> (packed);
  const int32_t *bases = reinterpret_cast<const int32_t *> (base + L.bases);
  const int32_t *raw = reinterpret_cast<const int32_t *> (base + L.raw);
  for (int y = 0; y < height; y++)
    for (int ci = 0; ci < L.chunks_per_row; ci++) {
      const size_t c = (size_t) y * L.chunks_per_row + ci;
      const int x0 = ci * CHUNK, n = (width - x0 < CHUNK) ? width - x0 : CHUNK;
      int32_t *out = index + (size_t) y * width + x0;
      if (bases[c] < 0) {
        B200VF_REQUIRE (~bases[c] < fh.n_raw, B200VF_E_INVAL, "gt_unpack_index: corrupt raw slot");
        memcpy (out, raw + (size_t) (~bases[c]) * CHUNK, (size_t) n * 4);
        continue;
      }
      int tx = bases[c] & 0xffff, ty = bases[c] >> 16;
      const uint8_t *st = base + L.steps + c * CHUNK;
      for (int i = 0; i < n; i++) {
        tx += (st[i] & 15) - 8; ty += (st[i] >> 4) - 8;
        out[i] = ty * width + tx;
      }
    }
  return B200VF_OK;
}

B200VF_API int b200vf_remap_packed (b200vf_ctx *ctx, const uint8_t *d_src, uint8_t *d_dst, const void *d_packed,
    int width, int height, size_t frame_stride, int nframes, uint32_t fill, void *stream)
{
  B200VF_REQUIRE (ctx && d_src && d_dst && d_packed && width > 0 && height > 0 && nframes > 0, B200VF_E_INVAL, "remap_packed: bad argument");
  B200VF_REQUIRE (width <= 32767 && height <= 32767, B200VF_E_UNSUPPORTED, "remap_packed: %dx%d", width, height);
  B200VF_REQUIRE (frame_stride >= (size_t) 4 * width * height && frame_stride % 4 == 0 && ((uintptr_t) d_src) % 4 == 0 &&
      ((uintptr_t) d_dst) % 4 == 0 && ((uintptr_t) d_packed) % 16 == 0, B200VF_E_INVAL, "remap_packed: strides / alignment");
  B200VF_REQUIRE (d_src != d_dst, B200VF_E_INVAL, "remap_packed: in-place remap is not defined");
  cudaStream_t s = b200vf_stream (ctx, stream);
  const PackedLayout L = packed_layout (width, height);
  const uint8_t *base = static_cast<const uint8_t *> (d_packed);
  // blocks per SM and frame: 1 is best for a batch (sweep in profiles/: 20.2k fps at 8K; more resident warps only spread
  // the gathers); a single frame needs 8 to fill the SMs, as in b200vf_remap (0.29 -> of the HBM peak otherwise)
  int per_sm = nframes >= 8 ? 1 : nframes >= 4 ? 2 : 8;
  if (const char *e = getenv ("B200VF_REMAP_BLOCKS_PER_SM")) { int v = atoi (e); if (v >= 1 && v <= 16) per_sm = v; }   // tuning knob
  int gx = ctx->sm_count * per_sm;
  const size_t need = (L.n_chunks + 255) / 256;                           // 8 warps x 32 chunks per block and iteration
  if (need < (size_t) gx) gx = (int) need;
  dim3 grid (gx, nframes);
  if (width % CHUNK == 0) remap4_packed_kernel<true><<<grid, 256, 0, s>>> (reinterpret_cast<const uint32_t *> (d_src), reinterpret_cast<uint32_t *> (d_dst),
      reinterpret_cast<const int32_t *> (base + L.bases), reinterpret_cast<const uint2 *> (base + L.steps),
      reinterpret_cast<const int32_t *> (base + L.raw), width, L.chunks_per_row, (uint32_t) L.n_chunks, frame_stride / 4, fill);
  else remap4_packed_kernel<false><<<grid, 256, 0, s>>> (reinterpret_cast<const uint32_t *> (d_src), reinterpret_cast<uint32_t *> (d_dst),
      reinterpret_cast<const int32_t *> (base + L.bases), reinterpret_cast<const uint2 *> (base + L.steps),
      reinterpret_cast<const int32_t *> (base + L.raw), width, L.chunks_per_row, (uint32_t) L.n_chunks, frame_stride / 4, fill);
  return b200vf_launched (ctx, "remap4_packed");
}

B200VF_API int b200vf_remap (b200vf_ctx *ctx, const uint8_t *d_src, uint8_t *d_dst, const int32_t *d_index,
    int width, int height, int pixel_stride, int row_stride, size_t frame_stride, int nframes,
    uint32_t fill, void *stream)
{
  B200VF_REQUIRE (ctx && d_src && d_dst && d_index && width > 0 && height > 0 && nframes > 0, B200VF_E_INVAL, "remap: bad argument");
  B200VF_REQUIRE (pixel_stride >= 1 && pixel_stride <= 4, B200VF_E_UNSUPPORTED, "remap: pixel stride %d", pixel_stride);
  B200VF_REQUIRE (row_stride >= pixel_stride * width && frame_stride >= (size_t) row_stride * height, B200VF_E_INVAL, "remap: strides");
  B200VF_REQUIRE (d_src != d_dst, B200VF_E_INVAL, "remap: in-place remap is not defined (the reference is out of place)");
  cudaStream_t s = b200vf_stream (ctx, stream);
  const size_t npix = (size_t) width * height;
  bool fast = pixel_stride == 4 && row_stride == 4 * width && ((uintptr_t) d_src) % 4 == 0 && ((uintptr_t) d_dst) % 4 == 0 &&
      ((uintptr_t) d_index) % 4 == 0 && frame_stride % 4 == 0;
  if (fast) {
    // Blocks per SM over the WHOLE launch (grid.x * nframes) decide how much latency the dependent index -> gather ->
    // store chain can hide: ~20 per SM measured best for a batch of 11 frames (2 per SM and frame, sweep 1..8 in
    // profiles/); a single frame launched with 2 per SM left the SMs at 25 % occupancy and ran at 0.33 of the HBM
    // peak (profiles/r02_remap.md), so few frames get more blocks each (8 resident blocks of 256 threads fill an SM).
    int per_sm = nframes >= 8 ? 2 : nframes >= 4 ? 4 : 8;
    if (const char *e = getenv ("B200VF_REMAP_BLOCKS_PER_SM")) { int v = atoi (e); if (v >= 1 && v <= 16) per_sm = v; }   // tuning knob
    int gx = ctx->sm_count * per_sm;
    size_t need = (npix + 1023) / 1024;                                  // 8 warps x 128 pixels per block and iteration
    if (need < (size_t) gx) gx = (int) need;
    dim3 grid (gx, nframes);
    remap4_kernel<<<grid, 256, 0, s>>> (reinterpret_cast<const uint32_t *> (d_src), reinterpret_cast<uint32_t *> (d_dst),
        d_index, npix, frame_stride / 4, fill);
    return b200vf_launched (ctx, "remap4");
  }
  dim3 grid ((width + 255) / 256, height, nframes);
  remap_generic_kernel<<<grid, 256, 0, s>>> (d_src, d_dst, d_index, width, height, pixel_stride, row_stride, frame_stride, fill);
  return b200vf_launched (ctx, "remap_generic");
}
