// lut.cuh - lane-replicated 256-entry word tables in shared memory.
// Entry v of lane l sits at word v*32+l, so the 32 lookups of a warp always hit 32
// different banks whatever the data (no bank conflicts for random pixels), and one
// 32-bit word carries up to four 8-bit tables, so a lookup is a single LDS.32.
#pragma once
#include "common.cuh"

struct PackedLut { uint32_t w[256]; };       // travels as a __grid_constant__ kernel parameter (1 KB)
constexpr int LUT_SMEM_BYTES = 256 * 32 * 4;

#ifdef __CUDACC__
__device__ __forceinline__ void lut_fill (uint32_t *tab, const PackedLut &lut) {
  const int nthreads = blockDim.x * blockDim.y;
  for (int i = threadIdx.y * blockDim.x + threadIdx.x; i < 256 * 32; i += nthreads) tab[i] = lut.w[i >> 5];
  __syncthreads ();
}
// per-byte-position LUT of one 4-byte pixel: byte c of w[v] = lut[c][v]
__device__ __forceinline__ uint32_t lut_px (const uint32_t *tab_lane, uint32_t in) {
  uint32_t w0 = tab_lane[(in & 0xff) << 5];
  uint32_t w1 = tab_lane[((in >> 8) & 0xff) << 5];
  uint32_t w2 = tab_lane[((in >> 16) & 0xff) << 5];
  uint32_t w3 = tab_lane[(in >> 24) << 5];
  uint32_t lo = PRMT (w0, w1, 0x7650);     // [w0.b0, w1.b1, .., ..]
  uint32_t hi = PRMT (w2, w3, 0x7210);     // [.., .., w2.b2, w3.b3]
  return PRMT (lo, hi, 0x7610);
}
#endif

static inline void pack_lut4 (const uint8_t lut[4][256], PackedLut &p) {
  for (int v = 0; v < 256; v++)
    p.w[v] = (uint32_t) lut[0][v] | ((uint32_t) lut[1][v] << 8) | ((uint32_t) lut[2][v] << 16) | ((uint32_t) lut[3][v] << 24);
}
