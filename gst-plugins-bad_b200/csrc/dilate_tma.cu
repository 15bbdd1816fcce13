// dilate_tma.cu - dilate / erode (gaudieffects) fed by TMA tiles: the sm_100a fast path of b200vf_dilate.
//
// The direct kernel (colorops.cu) keeps one 128-bit load in flight per thread while it marches down its
// strip and fetches the left/right halo pixels of each warp row with two extra 4-byte loads: it is bound by
// memory latency at 0.63 of the HBM roofline (profiles/). Here the bytes arrive the way they do in
// bayer_tma.cu: a persistent grid, one producer warp issuing cp.async.bulk.tensor boxes of
// (128 + 8) x (32 + 1) pixels - the tile, the row under it, and 4 pixels of margin left and right so
// that the box starts 16-byte aligned and holds both horizontal neighbours - into a 4-stage shared-memory
// ring (mbarrier full/empty pairs, tiles drawn from a global counter), and 8 consumer warps that read the
// ring with LDS.128 (+ two LDS.32 for the neighbours of the lane's 4 pixels) and write 128-bit streaming
// stores. Out-of-frame parts of a box arrive zero-filled and are never used: at the frame edges the
// neighbour is the pixel itself (gstdilate.c:283-305), applied by predicates.
#include "dilate.cuh"
#include "tma.cuh"
#include <stdlib.h>

namespace {

constexpr int TW = 128, TH = 32;            // output tile
constexpr int BW = TW + 8, BH = TH + 1;     // box in pixels: margin 4 left / 4 right, one row below
constexpr int STAGE_BYTES = ((BW * BH * 4 + 127) / 128) * 128;
constexpr int STAGES = 4;
constexpr int CONSUMERS = 8;                // warps; each takes TH / 8 = 4 rows of the tile
constexpr int ROWS_PER_WARP = TH / CONSUMERS;

struct DilParams {
  uint8_t *dst;
  size_t frame_stride;
  int width, height, rows_out, nframes;
  int tiles_x, tiles_y;
};

template <bool ERODE>
__global__ void __launch_bounds__ (CONSUMERS * 32 + 32)
dilate_tma_kernel (const __grid_constant__ CUtensorMap src_map, const DilParams p, unsigned int *tile_counter)
{
  extern __shared__ __align__ (128) uint8_t smem_raw[];
  __shared__ __align__ (8) uint64_t full[STAGES];
  __shared__ __align__ (8) uint64_t empty[STAGES];
  __shared__ int tile_of[STAGES];
  const int ntiles = p.tiles_x * p.tiles_y * p.nframes;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < STAGES; s++) { mbar_init (&full[s], 1); mbar_init (&empty[s], CONSUMERS); }
    mbar_fence_init ();
  }
  __syncthreads ();

  if (warp == CONSUMERS) {                  // ---------------------------------------- producer warp
    if (lane == 0) {
      for (int k = 0;; k++) {
        const int s = k % STAGES;
        if (k >= STAGES) mbar_wait (&empty[s], ((k / STAGES) - 1) & 1);
        const int t = (int) atomicAdd (tile_counter, 1u);
        tile_of[s] = t < ntiles ? t : -1;
        if (t >= ntiles) { mbar_arrive (&full[s]); break; }
        const int per_frame = p.tiles_x * p.tiles_y;
        const int f = t / per_frame, r = t - f * per_frame, ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
        mbar_expect_tx (&full[s], BW * BH * 4);
        tma_load_3d (smem_raw + s * STAGE_BYTES, &src_map, &full[s], tx * TW - 4, ty * TH, f);
      }
    }
    return;
  }

  // ------------------------------------------------------------------------------ consumer warps
  const int w = p.width;
  for (int k = 0;; k++) {
    const int s = k % STAGES;
    mbar_wait (&full[s], (k / STAGES) & 1);
    const int t = tile_of[s];
    if (t < 0) break;
    const int per_frame = p.tiles_x * p.tiles_y;
    const int f = t / per_frame, r = t - f * per_frame, ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
    const int x0 = tx * TW + lane * 4;                      // this lane's 4 pixels
    const int j0 = ty * TH + warp * ROWS_PER_WARP;
    const bool active = x0 < w;
    const bool first_col = x0 == 0, last_word = x0 + 4 >= w;
    // box pixel (bx, by) = frame pixel (tx*TW - 4 + bx, ty*TH + by)
    const uint32_t *box = reinterpret_cast<const uint32_t *> (smem_raw + s * STAGE_BYTES) + (warp * ROWS_PER_WARP) * BW + 4 + lane * 4;
    uint8_t *o = p.dst + (size_t) f * p.frame_stride + ((size_t) j0 * w + x0) * 4;
    uint4 cur = *reinterpret_cast<const uint4 *> (box);
    uint32_t cl[4] = { dil_lum (cur.x), dil_lum (cur.y), dil_lum (cur.z), dil_lum (cur.w) };
#pragma unroll
    for (int i = 0; i < ROWS_PER_WARP; i++) {
      const int j = j0 + i;
      if (j >= p.rows_out) break;                           // warp-uniform
      const uint32_t *rp = box + i * BW;
      uint4 nxt = *reinterpret_cast<const uint4 *> (rp + BW);
      if (j + 1 >= p.height) nxt = cur;                     // no row below: the pixel itself
      uint32_t left_in = rp[-1], right_in = rp[4];
      if (first_col) left_in = cur.x;                       // left of column 0 / right of the last column: itself
      if (last_word) right_in = cur.w;
      const uint32_t pv[4] = { cur.x, cur.y, cur.z, cur.w };
      const uint32_t dn[4] = { nxt.x, nxt.y, nxt.z, nxt.w };
      const uint32_t nl[4] = { dil_lum (nxt.x), dil_lum (nxt.y), dil_lum (nxt.z), dil_lum (nxt.w) };
      const uint32_t ll = dil_lum (left_in), rl = dil_lum (right_in);
      uint32_t ov[4];
#pragma unroll
      for (int q = 0; q < 4; q++) {
        uint32_t best = pv[q], bl = cl[q];
        dil_pick_t<ERODE> (best, bl, dn[q], nl[q]);                                          // down
        dil_pick_t<ERODE> (best, bl, (q < 3) ? pv[q + 1] : right_in, (q < 3) ? cl[q + 1] : rl);   // right
        dil_pick_t<ERODE> (best, bl, (q > 0) ? pv[q - 1] : left_in, (q > 0) ? cl[q - 1] : ll);    // left (`up` is dead code)
        ov[q] = best;
      }
      if (active) st_stream_v4 (o, make_uint4 (ov[0], ov[1], ov[2], ov[3]));
      o += (size_t) w * 4;
      cur = nxt;
#pragma unroll
      for (int q = 0; q < 4; q++) cl[q] = nl[q];
    }
    __syncwarp ();
    if (lane == 0) mbar_arrive (&empty[s]);
  }
}

}  // namespace

int b200vf_dilate_tma (b200vf_ctx *ctx, const uint8_t *d_src, uint8_t *d_dst, int width, int height, int rows_out,
    size_t frame_stride, int nframes, int erode, cudaStream_t s)
{
  CUtensorMap map;
  int rc = b200vf_encode_u32_3d (ctx, &map, d_src, (uint64_t) width, (uint64_t) height, (uint64_t) nframes,
      (uint64_t) width * 4, (uint64_t) frame_stride, BW, BH);
  if (rc) return rc;
  DilParams p;
  p.dst = d_dst; p.frame_stride = frame_stride; p.width = width; p.height = height; p.rows_out = rows_out; p.nframes = nframes;
  p.tiles_x = (width + TW - 1) / TW;
  p.tiles_y = (rows_out + TH - 1) / TH;
  const int smem = STAGES * STAGE_BYTES;
  {
    int rc;
    if ((rc = b200vf_func_smem (ctx, (const void *) dilate_tma_kernel<true>, smem)) ||
        (rc = b200vf_func_smem (ctx, (const void *) dilate_tma_kernel<false>, smem))) return rc;
  }
  const int ntiles = p.tiles_x * p.tiles_y * nframes;
  int per_sm = 2;                             // as for bayer2rgb_tma: more CTAs only add DRAM page conflicts
  if (const char *e = getenv ("B200VF_DILATE_CTAS_PER_SM")) { int v = atoi (e); if (v >= 1 && v <= 3) per_sm = v; }   // tuning knob
  int grid = ctx->sm_count * per_sm;
  if (grid > ntiles) grid = ntiles;
  unsigned int *counter = nullptr;
  rc = b200vf_next_tile_counter (ctx, s, &counter);
  if (rc) return rc;
  if (erode) dilate_tma_kernel<true><<<grid, CONSUMERS * 32 + 32, smem, s>>> (map, p, counter);
  else dilate_tma_kernel<false><<<grid, CONSUMERS * 32 + 32, smem, s>>> (map, p, counter);
  return b200vf_launched (ctx, "dilate_tma");
}
