// diffuse.cu - the one geometrictransform element without a precalculated map (gst/geometrictransform/gstdiffuse.c).
//
// Reference: gst_diffuse_init clears precalc_map (:228), so gst_geometric_transform_transform_frame
// (gstgeometrictransform.c:272-287) calls diffuse_map (:167-187) for every pixel of EVERY frame, in raster order:
//   angle = g_random_int_range (0, 256); distance = g_random_double ();
//   in = (x + distance * sin_table[angle], y + distance * cos_table[angle])
// with sin_table[i] = scale * sin (2 pi i / 256) built once by diffuse_prepare (:151-165), then do_map (:167-207).
// The draws come from GLib's global Mersenne twister seeded from the OS: a frame of the reference cannot be reproduced
// by anyone, the reference included (parity unpinned by construction). What CAN be pinned is everything
// around the draws, and that is what this kernel keeps: the fp64 coordinate arithmetic (separately rounded multiply
// and add: this file is compiled with --fmad=false), do_map's policy / truncation / bounds test, the cleared frame.
//
// The draws themselves are a counter-based generator - a function of (seed, frame number, pixel number) - so that
// every thread computes its own without state, a frame can be recomputed anywhere (b200vf_diffuse_draw is the same
// function on the host; tests rebuild whole frames from it and push them through the reference's do_map), and a
// row-sharded frame draws the same numbers whichever GPU owns the row. Statistics of the draws (uniform angle over
// 256 values, uniform distance in [0, 1) with 52 bits): tests/test_diffuse_gpu.py.
#include "common.cuh"
#include "gt_resolve.cuh"

namespace {

// splitmix64's finaliser over a Weyl sequence of the counter: 64 well-mixed bits per (seed, frame, pixel)
__host__ __device__ __forceinline__ uint64_t diffuse_bits (uint64_t seed, uint64_t frame, uint64_t pixel) {
  uint64_t z = seed + 0x9E3779B97F4A7C15ull * ((frame << 32) + pixel + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ int diffuse_angle (uint64_t z) { return (int) (z >> 56); }                       // g_random_int_range (0, 256)
__host__ __device__ __forceinline__ double diffuse_distance (uint64_t z) {                                           // g_random_double (): [0, 1)
  // bits 4..55 as the mantissa of a double in [1, 2), minus 1: k / 2^52, exact, and no 64-bit int -> double conversion
  const uint64_t bits = 0x3FF0000000000000ull | ((z >> 4) & 0x000FFFFFFFFFFFFFull);
  double d;
  memcpy (&d, &bits, sizeof d);
  return d - 1.0;
}

struct DiffuseParams {
  double sin_table[256], cos_table[256];       // scale * sin / cos (2 pi i / 256), from the host's libm
  int width, height, ps, row_stride, off_edge, first_row, full_height;
  size_t src_frame_stride, dst_frame_stride;
  uint32_t fill;
  uint64_t seed, first_frame;
};

// A block takes DPX consecutive pixels of one output row, thread t the pixels t, 256 + t, ... (coalesced stores);
// rows [first_row, first_row + height) of a full_height frame (row shards draw the numbers of their global pixel
// positions); the source is the whole frame. The displacement tables are staged in shared memory: indexed by a random
// angle they would serialise 32-fold in the constant bank (first version: 0.017 of the HBM peak).
constexpr int DPX = 1024;
template <bool WORDS> __global__ void __launch_bounds__ (256)
diffuse_kernel (const __grid_constant__ DiffuseParams p, const uint8_t *__restrict__ src, uint8_t *__restrict__ dst)
{
  __shared__ double2 s_tab[256];                 // (sin, cos): one 128-bit lookup per pixel
  s_tab[threadIdx.x] = make_double2 (p.sin_table[threadIdx.x], p.cos_table[threadIdx.x]);
  __syncthreads ();
  const uint8_t *s = src + (size_t) blockIdx.z * p.src_frame_stride;
  const uint64_t frame = p.first_frame + blockIdx.z;
  const int x0 = blockIdx.x * DPX + threadIdx.x;
  // the block walks down its column of the frame: the 4 KB of tables are staged once, not once per 1024 pixels
  // (indexed per lane, the constant bank serialises: the staging was 20 % of the warp time with one row per block)
  for (int yl = blockIdx.y; yl < p.height; yl += gridDim.y) {
    const int y = p.first_row + yl;
    uint8_t *d = dst + (size_t) blockIdx.z * p.dst_frame_stride + (size_t) yl * p.row_stride;
    const double yd = y;
    double xd = x0;                              // one int -> double conversion per thread and row, then exact additions
#pragma unroll
    for (int j = 0; j < DPX / 256; j++, xd += 256.0) {
      const int x = x0 + j * 256;
      if (x >= p.width) break;
      const uint64_t z = diffuse_bits (p.seed, frame, (uint64_t) y * p.width + x);
      const double2 sc = s_tab[diffuse_angle (z)];
      const double distance = diffuse_distance (z);
      const double in_x = xd + distance * sc.x;
      const double in_y = yd + distance * sc.y;
      int tx, ty;
      const bool mapped = resolve_xy (in_x, in_y, p.width, p.full_height, p.off_edge, tx, ty);
      if (WORDS) {
        const uint32_t v = mapped ? __ldg (reinterpret_cast<const uint32_t *> (s + (size_t) ty * p.row_stride) + tx) : p.fill;
        st_stream_u32 (d + (size_t) x * 4, v);
      } else {
        uint8_t *o = d + (size_t) x * p.ps;
        if (mapped) {
          const uint8_t *in = s + (size_t) ty * p.row_stride + (size_t) tx * p.ps;
          for (int b = 0; b < p.ps; b++) o[b] = in[b];
        } else {
          for (int b = 0; b < p.ps; b++) o[b] = (uint8_t) (p.fill >> (8 * ((x * p.ps + b) & 3)));
        }
      }
    }
    // row padding belongs to the cleared frame (memset covers map[0].size)
    if (blockIdx.x == 0)
      for (int b = p.width * p.ps + threadIdx.x; b < p.row_stride; b += blockDim.x) d[b] = (uint8_t) (p.fill >> (8 * (b & 3)));
  }
}

}  // namespace

B200VF_API void b200vf_diffuse_draw (uint64_t seed, uint64_t frame, uint64_t pixel, int *angle, double *distance) {
  const uint64_t z = diffuse_bits (seed, frame, pixel);
  if (angle) *angle = diffuse_angle (z);
  if (distance) *distance = diffuse_distance (z);
}

B200VF_API int b200vf_diffuse (b200vf_ctx *ctx, const uint8_t *d_src, uint8_t *d_dst, int width, int height, int first_row,
    int full_height, int pixel_stride, int row_stride, size_t src_frame_stride, size_t dst_frame_stride, int nframes, const double *sin_table,
    const double *cos_table, int off_edge, uint32_t fill, uint64_t seed, uint64_t first_frame, void *stream)
{
  B200VF_REQUIRE (ctx && d_src && d_dst && sin_table && cos_table && width > 0 && height > 0 && nframes > 0, B200VF_E_INVAL, "diffuse: bad argument");
  B200VF_REQUIRE (first_row >= 0 && full_height >= first_row + height, B200VF_E_INVAL, "diffuse: rows [%d, %d) of %d", first_row, first_row + height, full_height);
  B200VF_REQUIRE (pixel_stride >= 1 && pixel_stride <= 4, B200VF_E_UNSUPPORTED, "diffuse: pixel stride %d", pixel_stride);
  B200VF_REQUIRE (row_stride >= pixel_stride * width && dst_frame_stride >= (size_t) row_stride * height &&
      src_frame_stride >= (size_t) row_stride * full_height, B200VF_E_INVAL, "diffuse: strides");
  B200VF_REQUIRE (off_edge >= 0 && off_edge <= 2, B200VF_E_PROPERTY, "diffuse: off-edge-pixels %d", off_edge);
  B200VF_REQUIRE (d_src != d_dst, B200VF_E_INVAL, "diffuse: in-place remap is not defined (the reference is out of place)");
  B200VF_REQUIRE (height <= 65535 && nframes <= 65535 && (long long) width * full_height < 0x7fffffffll, B200VF_E_INVAL, "diffuse: frame too large");
  cudaStream_t s = b200vf_stream (ctx, stream);
  DiffuseParams p;
  memcpy (p.sin_table, sin_table, sizeof p.sin_table);
  memcpy (p.cos_table, cos_table, sizeof p.cos_table);
  p.width = width; p.height = height; p.ps = pixel_stride; p.row_stride = row_stride; p.off_edge = off_edge;
  p.first_row = first_row; p.full_height = full_height;
  p.src_frame_stride = src_frame_stride; p.dst_frame_stride = dst_frame_stride; p.fill = fill; p.seed = seed; p.first_frame = first_frame;
  const int gx = (width + DPX - 1) / DPX;
  long long gy = (long long) ctx->sm_count * 10 / ((long long) gx * nframes);          // ~10 blocks per SM over the whole launch
  gy = gy < 1 ? 1 : (gy > height ? height : gy);
  dim3 grid (gx, (unsigned) gy, nframes);
  const bool words = pixel_stride == 4 && row_stride % 4 == 0 && src_frame_stride % 4 == 0 && dst_frame_stride % 4 == 0 && ((uintptr_t) d_src) % 4 == 0 && ((uintptr_t) d_dst) % 4 == 0;
  if (words) diffuse_kernel<true><<<grid, 256, 0, s>>> (p, d_src, d_dst);
  else diffuse_kernel<false><<<grid, 256, 0, s>>> (p, d_src, d_dst);
  return b200vf_launched (ctx, "diffuse");
}

// diffuse_prepare (gstdiffuse.c:151-165): the displacement tables, with the host's libm
B200VF_API int b200vf_diffuse_tables (double scale, double *sin_table, double *cos_table) {
  B200VF_REQUIRE (sin_table && cos_table, B200VF_E_INVAL, "diffuse_tables: NULL argument");
  const double pi = 3.1415926535897932384626433832795028841971693993751;   // G_PI
  for (int i = 0; i < 256; i++) {
    double angle = (pi * 2 * i) / 256.0;
    sin_table[i] = scale * sin (angle);
    cos_table[i] = scale * cos (angle);
  }
  return B200VF_OK;
}
