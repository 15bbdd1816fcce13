// dilate.cuh - the per-pixel arithmetic of dilate shared by colorops.cu (direct kernels) and dilate_tma.cu.
// gst/gaudieffects/gstdilate.c:258-345: best of {self, down, right, left} by luminance
// 90 r + 115 g + 51 b, strict compare in that order (`up` is dead code, :291-294).
#pragma once
#include "common.cuh"

static __device__ __forceinline__ uint32_t dil_lum (uint32_t in) {
  return __dp4a (in, 0x005a7333u, 0u);          // 51*b0 + 115*b1 + 90*b2 (+ 0*x): one IDP4A
}
template <bool ERODE>
static __device__ __forceinline__ void dil_pick_t (uint32_t &best, uint32_t &bl, uint32_t cand, uint32_t cl) {
  const bool take = ERODE ? (cl < bl) : (cl > bl);
  best = take ? cand : best;
  bl = take ? cl : bl;
}

// TMA-fed variant (dilate_tma.cu): rows [0, rows_out) of frames whose rows [0, height) are in d_src;
// a row without a row under it in d_src (rows_out == height) uses itself as `down`.
int b200vf_dilate_tma (b200vf_ctx *ctx, const uint8_t *d_src, uint8_t *d_dst, int width, int height, int rows_out,
    size_t frame_stride, int nframes, int erode, cudaStream_t s);
