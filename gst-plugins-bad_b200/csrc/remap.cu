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

}  // namespace

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
