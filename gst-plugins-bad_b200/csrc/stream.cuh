// stream.cuh - TMA-fed streaming skeleton for the point ops (lut4, exclusion, coloreffects, chromahold).
//
// A point op reads 4 B and writes 4 B per pixel and does a handful of integer instructions in between;
// what decides its frame rate is how many bytes are in flight. The grid-stride kernels (pointops.cu,
// colorops.cu) keep 4 x 16 B per thread in registers and reach ~0.90 of the measured HBM peak; this
// skeleton moves the prefetch out of the register file, the way bayer_tma.cu / dilate_tma.cu do:
// a persistent grid, one producer thread issuing 1-D bulk copies (cp.async.bulk, 16 KB chunks of the
// flat pixel stream) into a 4-stage shared-memory ring guarded by full/empty mbarriers, chunks drawn
// from a global counter; 8 or 16 consumer warps read the ring with LDS.128, apply OP per pixel (OP's 256-entry
// word table sits lane-replicated in shared memory next to the ring) and write 128-bit streaming stores.
// In-place ops pass dst == src: a chunk is written back only after it has been copied out.
// Requires: src, dst 16-byte aligned, nbytes a multiple of 16 (callers peel heads/tails).
//
// OP: struct with  `void fill (uint32_t *tab) const`  (cooperative, all threads, ends with __syncthreads)
//                  `uint32_t apply (const uint32_t *tab_lane, uint32_t px) const`
#pragma once
#include "tma.cuh"
#include <stdlib.h>

constexpr int ST_CHUNK = 16384;             // bytes per stage
constexpr int ST_STAGES = 4;
constexpr int ST_TAB_BYTES = 256 * 32 * 4;  // lane-replicated table
constexpr int ST_SMEM = ST_TAB_BYTES + ST_STAGES * ST_CHUNK;

static __device__ __forceinline__ void bulk_load_1d (void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
  asm volatile ("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      :: "r"(smem_u32 (smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32 (bar)) : "memory");
}

// ST_CONSUMERS consumer warps (8 or 16) + 1 producer warp: light ops are fed by 8, ops with tens of
// instructions per pixel (chromahold's hue, lut4's four lookups) need 16 to keep up with the ring.
template <class OP, int ST_CONSUMERS>
__global__ void __launch_bounds__ (ST_CONSUMERS * 32 + 32)
stream_kernel (const uint8_t *src, uint8_t *dst, size_t nbytes, const __grid_constant__ OP op, unsigned int *chunk_counter)
{
  extern __shared__ __align__ (128) uint8_t st_smem[];
  uint32_t *tab = reinterpret_cast<uint32_t *> (st_smem);
  uint8_t *ring = st_smem + ST_TAB_BYTES;
  __shared__ __align__ (8) uint64_t full[ST_STAGES];
  __shared__ __align__ (8) uint64_t empty[ST_STAGES];
  __shared__ long long chunk_of[ST_STAGES];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < ST_STAGES; s++) { mbar_init (&full[s], 1); mbar_init (&empty[s], ST_CONSUMERS); }
    mbar_fence_init ();
  }
  op.fill (tab);                            // includes the CTA barrier that publishes the mbarriers
  const long long nchunks = (long long) ((nbytes + ST_CHUNK - 1) / ST_CHUNK);

  if (warp == ST_CONSUMERS) {               // ---------------------------------------- producer warp
    if (lane == 0) {
      for (int k = 0;; k++) {
        const int s = k % ST_STAGES;
        if (k >= ST_STAGES) mbar_wait (&empty[s], ((k / ST_STAGES) - 1) & 1);
        const long long c = (long long) atomicAdd (chunk_counter, 1u);
        chunk_of[s] = c < nchunks ? c : -1;
        if (c >= nchunks) { mbar_arrive (&full[s]); break; }
        const size_t off = (size_t) c * ST_CHUNK;
        const uint32_t bytes = (uint32_t) (nbytes - off < (size_t) ST_CHUNK ? nbytes - off : (size_t) ST_CHUNK);
        mbar_expect_tx (&full[s], bytes);
        bulk_load_1d (ring + s * ST_CHUNK, src + off, bytes, &full[s]);
      }
    }
    return;
  }

  const uint32_t *tl = tab + lane;          // ------------------------------------------ consumer warps
  for (int k = 0;; k++) {
    const int s = k % ST_STAGES;
    mbar_wait (&full[s], (k / ST_STAGES) & 1);
    const long long c = chunk_of[s];
    if (c < 0) break;
    const size_t off = (size_t) c * ST_CHUNK;
    const uint32_t bytes = (uint32_t) (nbytes - off < (size_t) ST_CHUNK ? nbytes - off : (size_t) ST_CHUNK);
    const uint8_t *in = ring + s * ST_CHUNK;
    uint8_t *out = dst + off;
    uint4 v[ST_CHUNK / (ST_CONSUMERS * 32 * 16)];
#pragma unroll
    for (int i = 0; i < ST_CHUNK / (ST_CONSUMERS * 32 * 16); i++) {
      const uint32_t o = (uint32_t) (i * ST_CONSUMERS * 32 + tid) * 16;
      if (o < bytes) v[i] = *reinterpret_cast<const uint4 *> (in + o);
    }
    __syncwarp ();
    if (lane == 0) mbar_arrive (&empty[s]); // the stage is in registers: hand it back before computing
#pragma unroll
    for (int i = 0; i < ST_CHUNK / (ST_CONSUMERS * 32 * 16); i++) {
      const uint32_t o = (uint32_t) (i * ST_CONSUMERS * 32 + tid) * 16;
      if (o < bytes) {
        uint4 r;
        r.x = op.apply (tl, v[i].x); r.y = op.apply (tl, v[i].y); r.z = op.apply (tl, v[i].z); r.w = op.apply (tl, v[i].w);
        st_stream_v4 (out + o, r);
      }
    }
  }
}

// 0 = launched (counted under `name`), else a status. `use` = false when the TMA path is switched off
// (ctx variant "direct" or B200VF_NO_STREAM_TMA): the caller then runs its grid-stride kernel.
static inline bool stream_enabled (const b200vf_ctx *ctx) { return ctx->variant != 1 && !getenv ("B200VF_NO_STREAM_TMA"); }

template <class OP>
int stream_launch (b200vf_ctx *ctx, const uint8_t *src, uint8_t *dst, size_t nbytes, const OP &op, cudaStream_t s, const char *name,
    int consumers = 8) {
  {
    int rc0;
    if ((rc0 = b200vf_func_smem (ctx, (const void *) stream_kernel<OP, 8>, ST_SMEM)) ||
        (rc0 = b200vf_func_smem (ctx, (const void *) stream_kernel<OP, 16>, ST_SMEM))) return rc0;
  }
  if (const char *e = getenv ("B200VF_STREAM_CONSUMERS")) { int v = atoi (e); if (v == 8 || v == 16) consumers = v; }   // tuning knob
  const size_t nchunks = (nbytes + ST_CHUNK - 1) / ST_CHUNK;
  int per_sm = 2;                           // 2 x 96 KB of shared memory per SM
  if (const char *e = getenv ("B200VF_STREAM_CTAS_PER_SM")) { int v = atoi (e); if (v >= 1 && v <= 2) per_sm = v; }   // tuning knob
  size_t grid = (size_t) ctx->sm_count * per_sm;
  if (grid > nchunks) grid = nchunks;
  unsigned int *counter = nullptr;
  int rc = b200vf_next_tile_counter (ctx, s, &counter);
  if (rc) return rc;
  if (consumers == 16) stream_kernel<OP, 16><<<(unsigned) grid, 16 * 32 + 32, ST_SMEM, s>>> (src, dst, nbytes, op, counter);
  else stream_kernel<OP, 8><<<(unsigned) grid, 8 * 32 + 32, ST_SMEM, s>>> (src, dst, nbytes, op, counter);
  return b200vf_launched (ctx, name);
}
