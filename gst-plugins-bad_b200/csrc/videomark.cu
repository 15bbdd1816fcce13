// videomark.cu - the rest of the videosignal plugin: simplevideomark / simplevideomarkdetect.
//
// simplevideomark draws a row of black / white boxes into the luma samples at the bottom left of the frame
// (gst_video_mark_yuv, gst/videosignal/gstsimplevideomark.c:348-462: `pattern-count` calibration boxes alternating
// black, white, then `pattern-data-count` boxes spelling `pattern-data`, most significant bit first);
// simplevideomarkdetect averages the same boxes (gst_video_detect_yuv, gstsimplevideomarkdetect.c:420-565),
// checks the calibration boxes against pattern-center +- pattern-sensitivity and reads the data bits back.
// The work is a few hundred samples per frame: what matters here is that the frame does not leave HBM for it - the
// mark is drawn in place by one small launch, the detector brings back one uint64 sum per box (72 bytes at the
// defaults) and decides on the host in the reference's double arithmetic.
// The box geometry (clipping against the frame, the early exits, the `continue`s that do not advance) is the
// reference's walk, replayed on the host by videomark_walk for both elements.
#include "common.cuh"
#include <string.h>

namespace {

constexpr int MAXB = B200VF_VIDEOMARK_MAX_BOXES;
struct MarkBoxes {
  int n;                         // boxes, in the order the reference visits them
  int n_calib;                   // the first n_calib belong to the calibration loop
  int calib_visited, data_visited;   // loop iterations that reached the brightness / draw call (detect: for the decision)
  long long off[MAXB];           // byte offset of the box's first sample from the first luma sample
  int w[MAXB], h[MAXB];          // samples per row, rows
  unsigned char color[MAXB];     // mark: 0 / 255
};

int calculate_pw (int pw, int x, int width) {                 // gstsimplevideomark.c:336-345
  if (x < 0) pw += x;
  else if ((x + pw) > width) pw = width - x;
  return pw;
}

// The walk both elements share. detect = 0: boxes as gst_video_mark_yuv draws them (clipped width draw_pw, colour);
// detect = 1: boxes as gst_video_detect_yuv averages them (the FULL pattern width even where the box is clipped -
// the reference reads on into the next row there, :476-479 - and a box is analysed before the clip test).
// Returns 0, or B200VF_E_UNSUPPORTED when more than MAXB boxes are visible.
int videomark_walk (const b200vf_videomark_params *P, int width, int height, int row_stride, int pixel_stride, int detect,
    uint64_t pattern_data, MarkBoxes *mb)
{
  memset (mb, 0, sizeof *mb);
  int pw = P->pattern_width, ph = P->pattern_height;
  const int pc = P->pattern_count, pdc = P->pattern_data_count;
  long long offset_calc = (long long) row_stride * (height - ph - P->bottom_offset) + (long long) pixel_stride * P->left_offset;
  int x = P->left_offset, y = height - ph - P->bottom_offset;
  const long long total = (long long) pc + pdc;
  // outside the video: nothing to draw / analyse (:381-386)
  if ((x + pw * total) < 0 || x > width || (y + height) < 0 || y > height) return B200VF_OK;
  if (offset_calc < 0) offset_calc = 0;
  if (y < 0) ph += y;
  else if ((y + ph) > height) ph = height - y;
  if (ph < 0) return B200VF_OK;
  long long d = offset_calc;
  auto add = [&] (int w, unsigned char color) -> bool {
    if (mb->n >= MAXB) return false;
    mb->off[mb->n] = d; mb->w[mb->n] = w; mb->h[mb->n] = ph; mb->color[mb->n] = color;
    mb->n++;
    return true;
  };
  bool stop = false;
  for (int i = 0; i < pc && !stop; i++) {
    if (detect) { if (!add (pw, 0)) return B200VF_E_UNSUPPORTED; mb->calib_visited++; }
    const int draw_pw = calculate_pw (pw, x, width);
    if (draw_pw < 0) continue;
    if (!detect) { if (!add (draw_pw, (i & 1) ? 255 : 0)) return B200VF_E_UNSUPPORTED; mb->calib_visited++; }
    d += (long long) pixel_stride * draw_pw;
    x += draw_pw;
    if ((x + pw * (total - i - 1)) < 0 || x >= width) stop = true;       // mark: return; detect: break out of THIS loop only
  }
  mb->n_calib = mb->n;
  if (stop && !detect) return B200VF_OK;                                   // gst_video_mark_yuv returns (:427-428)
  uint64_t shift = pdc > 0 ? (1ull << (pdc - 1)) : 0;
  for (int i = 0; i < pdc; i++) {
    if (detect) { if (!add (pw, 0)) return B200VF_E_UNSUPPORTED; mb->data_visited++; }
    const int draw_pw = calculate_pw (pw, x, width);
    if (draw_pw < 0) continue;
    if (!detect) { if (!add (draw_pw, (pattern_data & shift) ? 255 : 0)) return B200VF_E_UNSUPPORTED; mb->data_visited++; }
    shift >>= 1;
    d += (long long) pixel_stride * draw_pw;
    x += draw_pw;
    if ((x + pw * ((long long) pdc - i - 1)) < 0 || x >= width) break;
  }
  return B200VF_OK;
}

__global__ void __launch_bounds__ (128)
videomark_draw_kernel (uint8_t *luma, int pixel_stride, int row_stride, size_t frame_stride, long long plane_bytes,
    const __grid_constant__ MarkBoxes mb)
{
  const int b = blockIdx.x;
  uint8_t *d = luma + (size_t) blockIdx.y * frame_stride;
  const int w = mb.w[b], h = mb.h[b];
  for (int i = threadIdx.x; i < w * h; i += blockDim.x) {
    const long long o = mb.off[b] + (long long) (i / w) * row_stride + (long long) (i % w) * pixel_stride;
    if (o >= 0 && o < plane_bytes) d[o] = mb.color[b];
  }
}

__global__ void __launch_bounds__ (128)
videomark_sum_kernel (const uint8_t *luma, int pixel_stride, int row_stride, size_t frame_stride, long long plane_bytes,
    const __grid_constant__ MarkBoxes mb, unsigned long long *sums)
{
  const int b = blockIdx.x;
  const uint8_t *d = luma + (size_t) blockIdx.y * frame_stride;
  const int w = mb.w[b], h = mb.h[b];
  unsigned long long s = 0;
  for (long long i = threadIdx.x; i < (long long) w * h; i += blockDim.x) {
    const long long o = mb.off[b] + (i / w) * row_stride + (i % w) * pixel_stride;
    if (o >= 0 && o < plane_bytes) s += d[o];              // (the reference would read past the plane there: undefined)
  }
  __shared__ unsigned long long part[4];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync (0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
  __syncthreads ();
  if (threadIdx.x == 0) sums[(size_t) blockIdx.y * MAXB + b] = part[0] + part[1] + part[2] + part[3];
}

int check_params (const b200vf_videomark_params *P, int width, int height, int row_stride, int pixel_stride) {
  B200VF_REQUIRE (P && width > 0 && height > 0 && pixel_stride >= 1 && row_stride >= width * pixel_stride, B200VF_E_INVAL, "videomark: geometry");
  B200VF_REQUIRE (P->pattern_width >= 1 && P->pattern_height >= 1 && P->pattern_count >= 0 && P->pattern_data_count >= 0 &&
      P->left_offset >= 0 && P->bottom_offset >= 0, B200VF_E_PROPERTY, "videomark: property out of range");
  // the reference computes these products in int: beyond that its behaviour is undefined
  B200VF_REQUIRE ((long long) P->pattern_width * ((long long) P->pattern_count + P->pattern_data_count) < 0x7fffffffll, B200VF_E_UNSUPPORTED,
      "videomark: pattern-width x pattern count overflows int (undefined in the reference)");
  return B200VF_OK;
}

}  // namespace

B200VF_API int b200vf_videomark_draw (b200vf_ctx *ctx, uint8_t *d_luma, int pixel_stride, int row_stride, size_t frame_stride,
    int nframes, int width, int height, const b200vf_videomark_params *params, uint64_t pattern_data, void *stream)
{
  B200VF_REQUIRE (ctx && d_luma && nframes > 0, B200VF_E_INVAL, "videomark_draw: bad argument");
  int rc = check_params (params, width, height, row_stride, pixel_stride);
  if (rc) return rc;
  B200VF_REQUIRE (params->pattern_data_count <= 64, B200VF_E_PROPERTY, "videomark_draw: pattern-data-count %d > 64", params->pattern_data_count);
  MarkBoxes mb;
  rc = videomark_walk (params, width, height, row_stride, pixel_stride, 0, pattern_data, &mb);
  B200VF_REQUIRE (rc == B200VF_OK, rc, "videomark: more than %d boxes are visible", MAXB);
  if (mb.n == 0) return B200VF_OK;                           // "pattern is outside the video. Not drawing." (:383-385)
  cudaStream_t s = b200vf_stream (ctx, stream);
  videomark_draw_kernel<<<dim3 (mb.n, nframes), 128, 0, s>>> (d_luma, pixel_stride, row_stride, frame_stride,
      (long long) row_stride * height, mb);
  return b200vf_launched (ctx, "videomark_draw");
}

B200VF_API int b200vf_videomark_box_sums (b200vf_ctx *ctx, const uint8_t *d_luma, int pixel_stride, int row_stride, size_t frame_stride,
    int nframes, int width, int height, const b200vf_videomark_params *params, uint64_t *d_sums, int *n_boxes, void *stream)
{
  B200VF_REQUIRE (ctx && d_luma && d_sums && nframes > 0, B200VF_E_INVAL, "videomark_box_sums: bad argument");
  int rc = check_params (params, width, height, row_stride, pixel_stride);
  if (rc) return rc;
  MarkBoxes mb;
  rc = videomark_walk (params, width, height, row_stride, pixel_stride, 1, 0, &mb);
  B200VF_REQUIRE (rc == B200VF_OK, rc, "videomark: more than %d boxes are visible", MAXB);
  if (n_boxes) *n_boxes = mb.n;
  if (mb.n == 0) return B200VF_OK;
  cudaStream_t s = b200vf_stream (ctx, stream);
  videomark_sum_kernel<<<dim3 (mb.n, nframes), 128, 0, s>>> (d_luma, pixel_stride, row_stride, frame_stride,
      (long long) row_stride * height, mb, reinterpret_cast<unsigned long long *> (d_sums));
  return b200vf_launched (ctx, "videomark_box_sums");
}

// gst_video_detect_yuv's decisions (:476-560) on the box sums of ONE frame, in the reference's arithmetic:
// brightness = sum / (255.0 * pattern_width * pattern_height') as double. *in_pattern is the element's state;
// *message = 1 when the element would post its "GstSimpleVideoMarkDetect" message (have-pattern = *in_pattern after the
// call, data = *data).
B200VF_API int b200vf_videomark_detect_decide (const b200vf_videomark_params *params, int width, int height, int row_stride,
    int pixel_stride, const uint64_t *sums, double pattern_center, double pattern_sensitivity, int *in_pattern, int *message,
    uint64_t *data)
{
  B200VF_REQUIRE (sums && in_pattern && message && data, B200VF_E_INVAL, "videomark_detect_decide: NULL argument");
  int rc = check_params (params, width, height, row_stride, pixel_stride);
  if (rc) return rc;
  *message = 0; *data = 0;
  MarkBoxes mb;
  rc = videomark_walk (params, width, height, row_stride, pixel_stride, 1, 0, &mb);
  B200VF_REQUIRE (rc == B200VF_OK, rc, "videomark: more than %d boxes are visible", MAXB);
  // outside the video / negative height: "Not Analyzing", no message, state untouched (:449-466)
  {
    const int ph0 = params->pattern_height, y = height - ph0 - params->bottom_offset, x = params->left_offset;
    const long long total = (long long) params->pattern_count + params->pattern_data_count;
    if ((x + params->pattern_width * total) < 0 || x > width || (y + height) < 0 || y > height) return B200VF_OK;
    int ph = ph0;
    if (y < 0) ph += y; else if ((y + ph) > height) ph = height - y;
    if (ph < 0) return B200VF_OK;
  }
  int k = 0;
  for (int i = 0; i < mb.calib_visited; i++, k++) {
    const double brightness = sums[k] / (255.0 * mb.w[k] * mb.h[k]);
    bool wrong;
    if (i & 1) wrong = brightness < (pattern_center + pattern_sensitivity);       // odd boxes must be white
    else wrong = brightness > (pattern_center - pattern_sensitivity);             // even boxes must be black
    if (wrong) {                                                                   // no_pattern (:552-560)
      if (*in_pattern) { *in_pattern = 0; *message = 1; *data = 0; }
      return B200VF_OK;
    }
  }
  uint64_t pattern_data = 0;
  for (int i = 0; i < mb.data_visited; i++, k++) {
    const double brightness = sums[k] / (255.0 * mb.w[k] * mb.h[k]);
    pattern_data <<= 1;
    if (brightness > pattern_center) pattern_data |= 1;
  }
  *in_pattern = 1;
  *message = 1;
  *data = pattern_data;
  return B200VF_OK;
}
