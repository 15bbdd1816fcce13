// bayer_tma.cu - bayer2rgb fed by TMA 2-D tiles (the sm_100a fast path).
//
// Same algebra as bayer.cu (bayer.cuh); what changes is how bytes reach the SM:
//   * a CUtensorMap describes the mosaic as a 3-D tensor (x in 32-bit words, y, frame);
//   * a persistent grid of CTAs (a multiple of the SM count) walks the tile list;
//     one elected thread issues `cp.async.bulk.tensor.3d` for tiles STAGES ahead,
//     each landing a (TILE_W+32) x (TILE_H+2) byte box - the tile plus its 1-pixel
//     stencil halo, 16 B of margin left/right so every lane reads aligned words -
//     in a shared-memory ring, completion signalled on an mbarrier (no registers
//     hold in-flight loads, so the prefetch depth does not cost occupancy);
//   * 8 warps consume a tile: a warp marches down an 8-row strip of 128 pixels
//     keeping the 3 upsampled rows of the stencil in registers, one 128-bit
//     streaming store (4 RGBx pixels) per lane per row;
//   * out-of-bounds box rows/columns arrive zero-filled; the frame-edge rules of the
//     reference (top mirrors row 1, bottom reuses row h-4, left/right copies,
//     gstbayer2rgb.c:360-380,429-448) are applied by row re-indexing inside the box
//     and by the PRMT edge selectors, never by reading the zero fill.
// Needs: source base/pitch/frame pitch multiples of 16 B (TMA), destination
// 16-byte aligned. Everything else goes to bayer2rgb_direct.
#include "bayer.cuh"
#include "tma.cuh"
#include <stdlib.h>

namespace {

constexpr int TILE_W = 256;                 // output pixels per tile row
constexpr int TILE_H = 32;
constexpr int BOX_W = TILE_W + 32;          // bytes: 16 B margin each side (left halo byte at column 15)
constexpr int BOX_H = TILE_H + 2;
constexpr int STAGE_BYTES = ((BOX_W * BOX_H + 127) / 128) * 128;
constexpr int STAGES = 4;
constexpr int TMA_THREADS = 256;
constexpr int STRIP = 8;                    // rows per warp per tile (TILE_H / 4 strips)

struct TmaParams {
  uint8_t *dst;               // first OUTPUT row of this call (global row `row0`), frame 0
  size_t dst_frame_stride;
  int dst_stride;
  int width, full_height;     // frame geometry (edge rules use global rows)
  int row0, rows;             // output rows [row0, row0+rows): the whole frame, or this rank's row shard
  int buf_row0;               // global row held by tensor row 0 (row0-1 when a halo row precedes the shard)
  int nframes;
  int tiles_x, tiles_y;
  int first_is_gr;
};

__device__ __forceinline__ void tile_coords (int t, const TmaParams &p, int &f, int &ty, int &tx) {
  int per_frame = p.tiles_x * p.tiles_y;
  f = t / per_frame;
  int r = t - f * per_frame;
  ty = r / p.tiles_x;
  tx = r - ty * p.tiles_x;
}

// one Bayer row out of the staged box: this lane's word and its two neighbours (aligned LDS.32,
// constant offsets from `rp`), upsampled. XEDGE = the tile touches the left/right frame edge.
template <bool XEDGE>
__device__ __forceinline__ BayerRow box_row (const uint8_t *rp, uint32_t selL, uint32_t selR) {
  const uint32_t prev = *reinterpret_cast<const uint32_t *> (rp - 4);
  const uint32_t cur = *reinterpret_cast<const uint32_t *> (rp);
  const uint32_t next = *reinterpret_cast<const uint32_t *> (rp + 4);
  if (XEDGE) return bayer_upsample (prev, cur, next, selL, selR);
  return bayer_upsample_interior (prev, cur, next);
}

// Interior strip: all STRIP rows and their two stencil rows are plain rows of the box, so every
// LDS has a compile-time offset, the bg/gr role of a row is a compile-time constant (FIG = role of
// the strip's first row) and nothing is predicated (lanes right of the frame skip the strip).
template <int ORDER, int MODE, bool XEDGE, int FIG>
__device__ __forceinline__ void strip_fast (const uint8_t *rp /* box row of global row j0-1, this lane's column */,
    uint8_t *o, int dst_stride, uint32_t selL, uint32_t selR, const uint32_t *tl, uint32_t wts)
{
  BayerRow u = box_row<XEDGE> (rp, selL, selR);
  BayerRow c = box_row<XEDGE> (rp + BOX_W, selL, selR);
#pragma unroll
  for (int i = 0; i < STRIP; i += 2) {            // rows alternate roles: two rows per iteration, roles fixed
    {
      BayerRow d = box_row<XEDGE> (rp + (i + 2) * BOX_W, selL, selR);
      uint32_t R, G, B;
      bayer_merge_ct<FIG != 0> (u, c, d, R, G, B);
      uint4 px = bayer_epilogue<MODE> (bayer_pack<ORDER> (R, G, B, 0xffffffffu), tl, wts);
      st_stream_v4 (o, px);
      o += dst_stride;
      u = c;
      c = d;
    }
    {
      BayerRow d = box_row<XEDGE> (rp + (i + 3) * BOX_W, selL, selR);
      uint32_t R, G, B;
      bayer_merge_ct<FIG == 0> (u, c, d, R, G, B);
      uint4 px = bayer_epilogue<MODE> (bayer_pack<ORDER> (R, G, B, 0xffffffffu), tl, wts);
      st_stream_v4 (o, px);
      o += dst_stride;
      u = c;
      c = d;
    }
  }
}

// Warp-specialised persistent kernel. Warp 8 is the producer: it draws tile numbers from a global
// counter (dynamic scheduling: CTAs that run on a slower SM / die simply take fewer tiles), publishes
// them through shared memory and issues the TMA box loads up to STAGES tiles ahead, gated by per-stage
// `empty` mbarriers. Warps 0-7 are consumers: each waits on the stage's `full` mbarrier, processes its
// own 8-row x 128-px strip and arrives on `empty` - no CTA-wide barrier anywhere in the loop, so a warp
// that finishes early starts the next tile while its siblings are still storing.
template <int ORDER, int MODE>
__global__ void __launch_bounds__ (TMA_THREADS + 32)
bayer2rgb_tma_kernel (const __grid_constant__ CUtensorMap src_map, const TmaParams p,
    const __grid_constant__ BayerEpilogue epi, unsigned int *tile_counter)
{
  extern __shared__ __align__ (128) uint8_t smem_raw[];
  uint8_t *stage_base = smem_raw;
  uint32_t *epi_tab = reinterpret_cast<uint32_t *> (smem_raw + STAGES * STAGE_BYTES);
  if (MODE != 0) lut_fill (epi_tab, epi.table);
  const uint32_t *tl = epi_tab + (threadIdx.x & 31);
  __shared__ __align__ (8) uint64_t full[STAGES];
  __shared__ __align__ (8) uint64_t empty[STAGES];
  __shared__ int tile_of[STAGES];

  const int ntiles = p.tiles_x * p.tiles_y * p.nframes;
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int s = 0; s < STAGES; s++) { mbar_init (&full[s], 1); mbar_init (&empty[s], TMA_THREADS / 32); }
    mbar_fence_init ();
  }
  __syncthreads ();

  const int warp = tid >> 5, lane = tid & 31;

  if (warp == TMA_THREADS / 32) {
    // ------------------------------------------------------------------ producer warp
    if (lane == 0) {
      for (int k = 0;; k++) {
        const int s = k % STAGES;
        if (k >= STAGES) mbar_wait (&empty[s], ((k / STAGES) - 1) & 1);     // all 8 consumer warps left this stage
        const int t = (int) atomicAdd (tile_counter, 1u);
        tile_of[s] = t < ntiles ? t : -1;                                  // -1: no more work
        if (t >= ntiles) {
          mbar_arrive (&full[s]);
          break;
        }
        int f, ty, tx;
        tile_coords (t, p, f, ty, tx);
        mbar_expect_tx (&full[s], BOX_W * BOX_H);
        // x in 32-bit words: the box starts 16 B left of the tile and one row above it
        tma_load_3d (stage_base + s * STAGE_BYTES, &src_map, &full[s], tx * (TILE_W / 4) - 4,
            p.row0 + ty * TILE_H - 1 - p.buf_row0, f);
      }
    }
    return;
  }

  // -------------------------------------------------------------------- consumer warps
  const int half = warp & 1, strip = warp >> 1;
  const int xl = half * 128 + lane * 4;             // pixel x inside the tile
  const int h = p.full_height, w = p.width;
  const int row_end = p.row0 + p.rows;

  for (int k = 0;; k++) {
    const int s = k % STAGES;
    mbar_wait (&full[s], (k / STAGES) & 1);         // the tile number (and its box) have landed
    const int t = tile_of[s];
    if (t < 0) break;
    int f, ty, tx;
    tile_coords (t, p, f, ty, tx);
    const int x0 = tx * TILE_W + xl;
    const int y0 = p.row0 + ty * TILE_H;            // global row of the tile's first output row
    const int j0 = y0 + strip * STRIP;
    const bool active = x0 < w;
    const bool xedge = (tx == 0) || (tx == p.tiles_x - 1);
    const uint32_t selL = bayer_selL (x0);
    const uint32_t selR = bayer_selR (active && x0 + 4 >= w, 4);
    const uint8_t *box = stage_base + s * STAGE_BYTES + 16 + xl;     // box row r = global row y0 - 1 + r
    uint8_t *o = p.dst + (size_t) f * p.dst_frame_stride + (size_t) (j0 - p.row0) * p.dst_stride + (size_t) x0 * 4;
    const int fig = (p.first_is_gr ^ j0) & 1;       // role of the strip's first row: 0 = "bg", 1 = "gr"

    // rows j0-1 .. j0+STRIP are ordinary frame rows (no top mirror, no bottom rule, nothing clipped)
    const bool plain = (j0 >= 1) && (j0 + STRIP < h) && (j0 + STRIP <= row_end);
    if (plain) {
      const uint8_t *rp = box + (j0 - y0) * BOX_W;
      if (!active) {
      } else if (!xedge) {
        if (fig) strip_fast<ORDER, MODE, false, 1> (rp, o, p.dst_stride, selL, selR, tl, epi.luma_weights);
        else strip_fast<ORDER, MODE, false, 0> (rp, o, p.dst_stride, selL, selR, tl, epi.luma_weights);
      } else {
        if (fig) strip_fast<ORDER, MODE, true, 1> (rp, o, p.dst_stride, selL, selR, tl, epi.luma_weights);
        else strip_fast<ORDER, MODE, true, 0> (rp, o, p.dst_stride, selL, selR, tl, epi.luma_weights);
      }
    } else if (j0 < row_end) {
      // strips touching the top / bottom of the frame (or the end of the shard): rows re-indexed
      // per the reference's edge rules (gstbayer2rgb.c:429-448), everything else identical
      const int jend = min (j0 + STRIP, row_end);
      auto load_row = [&] (int g) { return box_row<true> (box + (g - (y0 - 1)) * BOX_W, selL, selR); };
      BayerRow u = load_row (j0 == 0 ? 1 : j0 - 1);
      BayerRow c = load_row (j0);
#pragma unroll 1
      for (int jj = j0; jj < jend; jj++) {
        const int gd = (jj + 1 < h) ? jj + 1 : (h >= 4 ? h - 4 : 1);
        BayerRow d = load_row (gd);
        uint32_t R, G, B;
        bayer_merge (u, c, d, ((jj ^ p.first_is_gr) & 1) != 0, R, G, B);
        uint4 px = bayer_pack<ORDER> (R, G, B, 0xffffffffu);
        px = bayer_epilogue<MODE> (px, tl, epi.luma_weights);
        if (active) st_stream_v4 (o, px);
        o += p.dst_stride;
        u = c;
        c = d;
      }
    }
    __syncwarp ();                                  // every lane is done reading stage s
    if (lane == 0) mbar_arrive (&empty[s]);
  }
}

typedef CUresult (*EncodeTiledFn) (CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode (b200vf_ctx *ctx) {
  if (!ctx->tma_encode) {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint ("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      fn = nullptr;
    cudaGetLastError ();
    ctx->tma_encode = fn;
  }
  return (EncodeTiledFn) ctx->tma_encode;
}

}  // namespace

// Shared by the other TMA users (gaussblur): a 3-D tensor of 32-bit words (x, y, frame).
int b200vf_encode_u32_3d (b200vf_ctx *ctx, CUtensorMap *map, const void *base, uint64_t words_x, uint64_t rows,
    uint64_t frames, uint64_t row_pitch_bytes, uint64_t frame_pitch_bytes, uint32_t box_x, uint32_t box_y) {
  EncodeTiledFn encode = get_encode (ctx);
  B200VF_REQUIRE (encode, B200VF_E_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t gdim[3] = { words_x, rows, frames };
  cuuint64_t gstride[2] = { row_pitch_bytes, frames > 1 ? frame_pitch_bytes : row_pitch_bytes * rows };
  cuuint32_t box[3] = { box_x, box_y, 1 };
  cuuint32_t estr[3] = { 1, 1, 1 };
  CUresult r = encode (map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, (void *) base, gdim, gstride, box, estr,
      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  B200VF_REQUIRE (r == CUDA_SUCCESS, B200VF_E_CUDA, "cuTensorMapEncodeTiled failed: %d", (int) r);
  return B200VF_OK;
}

namespace {

template <int ORDER, int MODE>
int launch_tma_mode (b200vf_ctx *ctx, const CUtensorMap &map, const TmaParams &p, const BayerEpilogue &epi, cudaStream_t s) {
  const int smem = STAGES * STAGE_BYTES + (MODE ? LUT_SMEM_BYTES : 0);
  if (int rc = b200vf_func_smem (ctx, (const void *) bayer2rgb_tma_kernel<ORDER, MODE>, smem)) return rc;
  int ntiles = p.tiles_x * p.tiles_y * p.nframes;
  int per_sm = 2;                              // CTAs per SM (measured sweep 1..6 in profiles/: 2 is best; more CTAs only add DRAM page conflicts)
  if (const char *e = getenv ("B200VF_TMA_CTAS_PER_SM")) { int v = atoi (e); if (v >= 1 && v <= 8) per_sm = v; }   // tuning knob
  int grid = ctx->sm_count * per_sm;
  if (grid > ntiles) grid = ntiles;
  unsigned int *counter = nullptr;
  int rc = b200vf_next_tile_counter (ctx, s, &counter);
  if (rc) return rc;
  bayer2rgb_tma_kernel<ORDER, MODE><<<grid, TMA_THREADS + 32, smem, s>>> (map, p, epi, counter);
  return b200vf_launched (ctx, MODE ? "bayer2rgb_tma_fused" : "bayer2rgb_tma");
}
template <int ORDER>
int launch_tma (b200vf_ctx *ctx, const CUtensorMap &map, const TmaParams &p, const BayerEpilogue &epi, cudaStream_t s) {
  switch (epi.mode) {
    case 0: return launch_tma_mode<ORDER, 0> (ctx, map, p, epi, s);
    case 1: return launch_tma_mode<ORDER, 1> (ctx, map, p, epi, s);
    default: return launch_tma_mode<ORDER, 2> (ctx, map, p, epi, s);
  }
}

}  // namespace

bool b200vf_bayer2rgb_tma_usable (b200vf_ctx *ctx, const uint8_t *d_src, int src_stride, size_t src_frame_stride,
    const uint8_t *d_dst, int dst_stride, size_t dst_frame_stride, int width, int full_height, int row0, int rows) {
  if (((uintptr_t) d_src) % 16 || src_stride % 16 || src_frame_stride % 16) return false;
  if (((uintptr_t) d_dst) % 16 || dst_stride % 16 || dst_frame_stride % 16) return false;
  if (width % 4) return false;                 // the last word must be a full word
  if (full_height < 4) return false;
  if (row0 + rows == full_height) {            // row h-4 must lie inside the last tile's box
    int rem = rows % TILE_H;
    if (rem == 1 || rem == 2) return false;
  }
  if (row0 > 0 && ((uintptr_t) (d_src - src_stride)) % 16) return false;   // tensor base = the halo row above
  return get_encode (ctx) != nullptr;
}

// d_src points at global row `row0` (the shard's first row; one halo row above it when row0 > 0
// and one below when the shard does not end the frame).
int b200vf_bayer2rgb_tma_launch (b200vf_ctx *ctx, const uint8_t *d_src, int src_stride, size_t src_frame_stride,
    uint8_t *d_dst, int dst_stride, size_t dst_frame_stride, int width, int full_height, int row0, int rows, int nframes,
    int order, int first_is_gr, const BayerEpilogue &epi, cudaStream_t s)
{
  EncodeTiledFn encode = get_encode (ctx);
  B200VF_REQUIRE (encode, B200VF_E_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  const int buf_row0 = row0 > 0 ? row0 - 1 : 0;
  const int buf_end = (row0 + rows < full_height) ? row0 + rows + 1 : full_height;
  const uint8_t *base = d_src - (size_t) (row0 - buf_row0) * src_stride;
  CUtensorMap map;
  // 3-D tensor of 32-bit words: x (width/4 words; bytes past `width` inside the pitch are never used), y, frame
  cuuint64_t gdim[3] = { (cuuint64_t) (width / 4), (cuuint64_t) (buf_end - buf_row0), (cuuint64_t) nframes };
  cuuint64_t gstride[2] = { (cuuint64_t) src_stride,
    (cuuint64_t) (nframes > 1 ? src_frame_stride : (size_t) src_stride * (buf_end - buf_row0)) };
  cuuint32_t box[3] = { BOX_W / 4, BOX_H, 1 };
  cuuint32_t estr[3] = { 1, 1, 1 };
  CUresult r = encode (&map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, (void *) base, gdim, gstride, box, estr,
      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  B200VF_REQUIRE (r == CUDA_SUCCESS, B200VF_E_CUDA, "cuTensorMapEncodeTiled failed: %d", (int) r);
  TmaParams p;
  p.dst = d_dst;
  p.dst_frame_stride = dst_frame_stride;
  p.dst_stride = dst_stride;
  p.width = width;
  p.full_height = full_height;
  p.row0 = row0;
  p.rows = rows;
  p.buf_row0 = buf_row0;
  p.nframes = nframes;
  p.tiles_x = (width + TILE_W - 1) / TILE_W;
  p.tiles_y = (rows + TILE_H - 1) / TILE_H;
  p.first_is_gr = first_is_gr;
  switch (order) {
    case 0: return launch_tma<0> (ctx, map, p, epi, s);
    case 1: return launch_tma<1> (ctx, map, p, epi, s);
    case 2: return launch_tma<2> (ctx, map, p, epi, s);
    default: return launch_tma<3> (ctx, map, p, epi, s);
  }
}
