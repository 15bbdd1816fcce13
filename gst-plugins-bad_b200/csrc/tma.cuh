// tma.cuh - mbarrier + TMA (cp.async.bulk.tensor) primitives shared by the tile-fed kernels
// (bayer_tma.cu, gaussblur.cu, dilate_tma.cu). sm_100a PTX, one CTA per cluster.
#pragma once
#include "common.cuh"
#include <cuda.h>

// A 3-D tensor of 32-bit words (x, y, frame) with zero fill outside; defined in bayer_tma.cu.
int b200vf_encode_u32_3d (b200vf_ctx *ctx, CUtensorMap *map, const void *base, uint64_t words_x, uint64_t rows,
    uint64_t frames, uint64_t row_pitch_bytes, uint64_t frame_pitch_bytes, uint32_t box_x, uint32_t box_y);
// A zeroed work counter for a dynamically scheduled kernel launched on stream s (core.cu).
int b200vf_next_tile_counter (b200vf_ctx *ctx, cudaStream_t s, unsigned int **out);

static __device__ __forceinline__ uint32_t smem_u32 (const void *p) { return (uint32_t) __cvta_generic_to_shared (p); }

static __device__ __forceinline__ void mbar_init (uint64_t *bar, int count) {
  asm volatile ("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32 (bar)), "r"(count));
}
static __device__ __forceinline__ void mbar_fence_init () { asm volatile ("fence.mbarrier_init.release.cluster;" ::: "memory"); }
static __device__ __forceinline__ void mbar_expect_tx (uint64_t *bar, uint32_t bytes) {
  asm volatile ("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32 (bar)), "r"(bytes) : "memory");
}
static __device__ __forceinline__ void mbar_arrive (uint64_t *bar) {
  asm volatile ("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32 (bar)) : "memory");
}
static __device__ __forceinline__ void mbar_wait (uint64_t *bar, uint32_t parity) {
  asm volatile (
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" :: "r"(smem_u32 (bar)), "r"(parity) : "memory");
}
// box at element coordinates (c0, c1, c2) -> shared memory; completion (byte count) on `bar`
static __device__ __forceinline__ void tma_load_3d (void *smem_dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2) {
  asm volatile (
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      :: "r"(smem_u32 (smem_dst)), "l"(map), "r"(smem_u32 (bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
