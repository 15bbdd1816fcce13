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
#include <string.h>

namespace {

__device__ __forceinline__ int4 ld_idx4 (const int32_t *p) {
  int4 r;
  asm volatile ("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
      : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

// 4-byte pixels, rows contiguous (row_stride == 4*width): 4 output pixels per thread
__global__ void __launch_bounds__ (256)
remap4_kernel (const uint32_t *__restrict__ src, uint32_t *__restrict__ dst, const int32_t *__restrict__ index,
    size_t npix, size_t frame_px, uint32_t fill)
{
  const uint32_t *s = src + (size_t) blockIdx.y * frame_px;
  uint32_t *d = dst + (size_t) blockIdx.y * frame_px;
  const size_t n4 = npix / 4;
  const size_t stride = (size_t) gridDim.x * blockDim.x;
  for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    int4 ix = ld_idx4 (index + 4 * i);
    uint4 o;
    o.x = ix.x >= 0 ? __ldg (s + ix.x) : fill;
    o.y = ix.y >= 0 ? __ldg (s + ix.y) : fill;
    o.z = ix.z >= 0 ? __ldg (s + ix.z) : fill;
    o.w = ix.w >= 0 ? __ldg (s + ix.w) : fill;
    st_stream_v4 (d + 4 * i, o);
  }
  if (blockIdx.x == 0 && threadIdx.x < (int) (npix - n4 * 4)) {
    size_t i = n4 * 4 + threadIdx.x;
    int ix = index[i];
    d[i] = ix >= 0 ? __ldg (s + ix) : fill;
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
// the next, so the table is stored as: per group of 128 consecutive pixels of one output row the source
// position (tx, ty) of the first pixel (2 x int16) and one byte per pixel holding the step from the previous
// pixel, (dtx + 8) | (dty + 8) << 4 with both steps in [-8, 7]: 1.06 B/px. Groups that cannot be coded
// (an ignored pixel, a discontinuity of the map) keep their 128 int32 entries in a side array.
// One warp decodes one group: 4 steps per lane, local sums, one warp scan of (sum_x + 65536 * sum_y).
struct PackedHead { int16_t tx0, ty0; int32_t raw_slot; };   // raw_slot < 0: step-coded

__global__ void __launch_bounds__ (256)
remap4_packed_kernel (const uint32_t *__restrict__ src, uint32_t *__restrict__ dst, const PackedHead *__restrict__ heads,
    const uint32_t *__restrict__ steps, const int32_t *__restrict__ raw, int w, int groups_per_row, int n_groups,
    size_t frame_px, uint32_t fill)
{
  const uint32_t *s = src + (size_t) blockIdx.y * frame_px;
  uint32_t *d = dst + (size_t) blockIdx.y * frame_px;
  const int lane = threadIdx.x & 31;
  const int warps = gridDim.x * (blockDim.x / 32);
  for (int g = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5); g < n_groups; g += warps) {
    const int y = g / groups_per_row, x0 = (g - y * groups_per_row) * 128 + lane * 4;
    const int2 hd = *reinterpret_cast<const int2 *> (heads + g);          // same address for the warp: one broadcast load
    int4 ix;
    if (hd.y >= 0) {
      ix = ld_idx4 (raw + ((size_t) hd.y * 128 + lane * 4));
    } else {
      const uint32_t st = ldg_u32 (steps + (size_t) g * 32 + lane);
      const int dx0 = (int) (st & 15) - 8, dy0 = (int) ((st >> 4) & 15) - 8;
      const int dx1 = (int) ((st >> 8) & 15) - 8, dy1 = (int) ((st >> 12) & 15) - 8;
      const int dx2 = (int) ((st >> 16) & 15) - 8, dy2 = (int) ((st >> 20) & 15) - 8;
      const int dx3 = (int) ((st >> 24) & 15) - 8, dy3 = (int) (st >> 28) - 8;
      const int sx1 = dx0, sx2 = sx1 + dx1, sx3 = sx2 + dx2, sx4 = sx3 + dx3;
      const int sy1 = dy0, sy2 = sy1 + dy1, sy3 = sy2 + dy2, sy4 = sy3 + dy3;
      const int tot = sx4 + sy4 * 65536;                                  // |sums| <= 128 * 8: both fit 16 bits, the sum is linear
      int v = tot;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync (0xffffffffu, v, o);
        if (lane >= o) v += t;
      }
      const int ex = v - tot;                                             // steps of the lanes before this one
      const int bx = (int) (int16_t) (ex & 0xffff);
      const int by = (ex - bx) >> 16;
      const int tx = (int) (int16_t) (hd.x & 0xffff) + bx, ty = (hd.x >> 16) + by;
      ix.x = (ty + sy1) * w + tx + sx1;
      ix.y = (ty + sy2) * w + tx + sx2;
      ix.z = (ty + sy3) * w + tx + sx3;
      ix.w = (ty + sy4) * w + tx + sx4;
    }
    if (x0 >= w) continue;
    uint4 o;
    const bool full = x0 + 3 < w;
    o.x = ix.x >= 0 ? __ldg (s + ix.x) : fill;
    o.y = (x0 + 1 < w) ? (ix.y >= 0 ? __ldg (s + ix.y) : fill) : 0u;
    o.z = (x0 + 2 < w) ? (ix.z >= 0 ? __ldg (s + ix.z) : fill) : 0u;
    o.w = full ? (ix.w >= 0 ? __ldg (s + ix.w) : fill) : 0u;
    uint32_t *out = d + (size_t) y * w + x0;
    if (full && (w & 3) == 0) st_stream_v4 (out, o);
    else {
      out[0] = o.x;
      if (x0 + 1 < w) out[1] = o.y;
      if (x0 + 2 < w) out[2] = o.z;
      if (full) out[3] = o.w;
    }
  }
}

struct PackedLayout { size_t heads, steps, raw, n_groups; int groups_per_row; };
PackedLayout packed_layout (int w, int h) {
  PackedLayout L;
  L.groups_per_row = (w + 127) / 128;
  L.n_groups = (size_t) L.groups_per_row * h;
  L.heads = 64;                                                           // after a 64-byte file header
  L.steps = L.heads + ((L.n_groups * sizeof (PackedHead) + 63) & ~(size_t) 63);
  L.raw = L.steps + L.n_groups * 128;
  return L;
}
struct PackedFileHeader { uint32_t magic, version; int32_t w, h, groups_per_row, n_raw; uint64_t total; };
const uint32_t PACKED_MAGIC = 0x50523242u;                                // "B2RP"

}  // namespace

B200VF_API size_t b200vf_gt_packed_bound (int width, int height) {
  if (width <= 0 || height <= 0) return 0;
  PackedLayout L = packed_layout (width, height);
  return L.raw + L.n_groups * 512;
}

B200VF_API int b200vf_gt_pack_index (const int32_t *index, int width, int height, void *packed, size_t capacity,
    size_t *used, size_t *raw_groups)
{
  B200VF_REQUIRE (index && packed && width > 0 && height > 0, B200VF_E_INVAL, "gt_pack_index: bad argument");
  B200VF_REQUIRE (width <= 32767 && height <= 32767, B200VF_E_UNSUPPORTED, "gt_pack_index: %dx%d exceeds 16-bit positions", width, height);
  const PackedLayout L = packed_layout (width, height);
  B200VF_REQUIRE (capacity >= L.raw, B200VF_E_INVAL, "gt_pack_index: capacity %zu < %zu", capacity, L.raw);
  uint8_t *base = static_cast<uint8_t *> (packed);
  PackedHead *heads = reinterpret_cast<PackedHead *> (base + L.heads);
  uint8_t *steps = base + L.steps;
  int32_t *raw = reinterpret_cast<int32_t *> (base + L.raw);
  memset (base, 0, L.raw);
  size_t n_raw = 0;
  for (int y = 0; y < height; y++) {
    for (int gi = 0; gi < L.groups_per_row; gi++) {
      const size_t g = (size_t) y * L.groups_per_row + gi;
      const int x0 = gi * 128, n = (width - x0 < 128) ? width - x0 : 128;
      const int32_t *ix = index + (size_t) y * width + x0;
      uint8_t *st = steps + g * 128;
      bool ok = ix[0] >= 0;
      int px = 0, py = 0;
      if (ok) { px = ix[0] % width; py = ix[0] / width; st[0] = 0x88; }
      for (int i = 1; ok && i < n; i++) {
        if (ix[i] < 0) { ok = false; break; }
        const int tx = ix[i] % width, ty = ix[i] / width;
        const int dx = tx - px, dy = ty - py;
        if (dx < -8 || dx > 7 || dy < -8 || dy > 7) { ok = false; break; }
        st[i] = (uint8_t) ((dx + 8) | ((dy + 8) << 4));
        px = tx; py = ty;
      }
      if (ok) {
        for (int i = n; i < 128; i++) st[i] = 0x88;
        heads[g].tx0 = (int16_t) (ix[0] % width); heads[g].ty0 = (int16_t) (ix[0] / width); heads[g].raw_slot = -1;
      } else {
        if (L.raw + (n_raw + 1) * 512 > capacity) {
          b200vf_set_error ("gt_pack_index: capacity %zu too small (b200vf_gt_packed_bound)", capacity);
          return B200VF_E_INVAL;
        }
        memset (st, 0x88, 128);
        int32_t *r = raw + n_raw * 128;
        for (int i = 0; i < 128; i++) r[i] = i < n ? ix[i] : -1;
        heads[g].tx0 = 0; heads[g].ty0 = 0; heads[g].raw_slot = (int32_t) n_raw;
        n_raw++;
      }
    }
  }
  PackedFileHeader fh;
  fh.magic = PACKED_MAGIC; fh.version = 1; fh.w = width; fh.h = height; fh.groups_per_row = L.groups_per_row;
  fh.n_raw = (int32_t) n_raw; fh.total = L.raw + n_raw * 512;
  memcpy (base, &fh, sizeof fh);
  if (used) *used = (size_t) fh.total;
  if (raw_groups) *raw_groups = n_raw;
  return B200VF_OK;
}

// host decoder of the packed table (tests, tooling): the inverse of b200vf_gt_pack_index
B200VF_API int b200vf_gt_unpack_index (const void *packed, size_t size, int width, int height, int32_t *index) {
  B200VF_REQUIRE (packed && index && width > 0 && height > 0 && size >= 64, B200VF_E_INVAL, "gt_unpack_index: bad argument");
  PackedFileHeader fh;
  memcpy (&fh, packed, sizeof fh);
  B200VF_REQUIRE (fh.magic == PACKED_MAGIC && fh.version == 1 && fh.w == width && fh.h == height && fh.total <= size,
      B200VF_E_INVAL, "gt_unpack_index: not a packed table for %dx%d", width, height);
  const PackedLayout L = packed_layout (width, height);
  const uint8_t *base = static_cast<const uint8_t *> (packed);
  const PackedHead *heads = reinterpret_cast<const PackedHead *> (base + L.heads);
  const int32_t *raw = reinterpret_cast<const int32_t *> (base + L.raw);
  for (int y = 0; y < height; y++)
    for (int gi = 0; gi < L.groups_per_row; gi++) {
      const size_t g = (size_t) y * L.groups_per_row + gi;
      const int x0 = gi * 128, n = (width - x0 < 128) ? width - x0 : 128;
      int32_t *out = index + (size_t) y * width + x0;
      if (heads[g].raw_slot >= 0) {
        B200VF_REQUIRE (heads[g].raw_slot < fh.n_raw, B200VF_E_INVAL, "gt_unpack_index: corrupt raw slot");
        memcpy (out, raw + (size_t) heads[g].raw_slot * 128, (size_t) n * 4);
        continue;
      }
      int tx = heads[g].tx0, ty = heads[g].ty0;
      const uint8_t *st = base + L.steps + g * 128;
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
  B200VF_REQUIRE (frame_stride >= (size_t) 4 * width * height && frame_stride % 16 == 0 && ((uintptr_t) d_src) % 4 == 0 &&
      ((uintptr_t) d_dst) % 16 == 0 && ((uintptr_t) d_packed) % 16 == 0, B200VF_E_INVAL, "remap_packed: strides / alignment");
  B200VF_REQUIRE (d_src != d_dst, B200VF_E_INVAL, "remap_packed: in-place remap is not defined");
  cudaStream_t s = b200vf_stream (ctx, stream);
  const PackedLayout L = packed_layout (width, height);
  const uint8_t *base = static_cast<const uint8_t *> (d_packed);
  int gx = ctx->sm_count * 8;
  const size_t need = (L.n_groups + 7) / 8;
  if (need < (size_t) gx) gx = (int) need;
  dim3 grid (gx, nframes);
  remap4_packed_kernel<<<grid, 256, 0, s>>> (reinterpret_cast<const uint32_t *> (d_src), reinterpret_cast<uint32_t *> (d_dst),
      reinterpret_cast<const PackedHead *> (base + L.heads), reinterpret_cast<const uint32_t *> (base + L.steps),
      reinterpret_cast<const int32_t *> (base + L.raw), width, L.groups_per_row, (int) L.n_groups, frame_stride / 4, fill);
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
  bool fast = pixel_stride == 4 && row_stride == 4 * width && ((uintptr_t) d_src) % 4 == 0 && ((uintptr_t) d_dst) % 16 == 0 &&
      ((uintptr_t) d_index) % 16 == 0 && frame_stride % 16 == 0;
  if (fast) {
    int gx = ctx->sm_count * 8;
    size_t need = (npix / 4 + 255) / 256;
    if (need < (size_t) gx) gx = need ? (int) need : 1;
    dim3 grid (gx, nframes);
    remap4_kernel<<<grid, 256, 0, s>>> (reinterpret_cast<const uint32_t *> (d_src), reinterpret_cast<uint32_t *> (d_dst),
        d_index, npix, frame_stride / 4, fill);
    return b200vf_launched (ctx, "remap4");
  }
  dim3 grid ((width + 255) / 256, height, nframes);
  remap_generic_kernel<<<grid, 256, 0, s>>> (d_src, d_dst, d_index, width, height, pixel_stride, row_stride, frame_stride, fill);
  return b200vf_launched (ctx, "remap_generic");
}
