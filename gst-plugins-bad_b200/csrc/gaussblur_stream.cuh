// gaussblur_stream.cuh - the streaming form of the exact gaussianblur (included by gaussblur.cu, inside its
// anonymous namespace, after the arithmetic helpers).
//
// Same arithmetic as gaussblur_kernel, operation for operation as far as any output byte can tell, with HALF the
// multiplies. make_gaussian_kernel (gstgaussblur.c:361-422) produces bitwise symmetric taps (k[i] == k[2C-i]: the
// taps are a function of (i - C)^2, normalised by one common sum), so the separately rounded product RN (x[s] * k[j])
// of sample s is the same number for the two outputs s + (C - j) and s - (C - j). A thread therefore walks ALONG the
// blur axis with a register ring of accumulators, one per output in flight: for every sample it forms the C + 1
// distinct products once (FMUL2, two channels per instruction) and adds them to the 2C + 1 outputs the sample
// contributes to (FFMA2 by an opaque 1.0 = a separately rounded add, see tapN), each output still receiving its
// addends in ascending tap order. 3C + 1 FMA-pipe instructions per sample and channel pair instead of 2 (2C + 1):
// 40 instead of 54 at sigma = 5 - the bytes are the reference's, the FP32 "roofline" of 16 T lane-operations per
// pixel is no longer a floor. (The fresh accumulator is SET to its first product instead of added to
// +0: identical except that -0 stays -0, which no later operation can turn into a different byte.)
//
// One CTA per SM, 16 warps, specialised:
//   warps 0-7    horizontal pass. Input by TMA: one box (32 rows x (128 + 2C) aligned u8x4 samples, zero-filled outside
//                the frame) per block into a 2-deep ring. lane = (row, channel pair), warp = 16 rows of one 32-pixel
//                segment; a thread streams over the 32 + 2C samples of its segment (fully unrolled: products that no
//                output of the segment needs are never formed), divides, and writes fp32 rows into a 2-deep ring of
//                32-row blocks;
//   warps 8-15   vertical pass: lane = (column, channel pair) of a 128-pixel strip; a thread streams DOWN the strip,
//                its ring lives in registers for the whole walk (no halo rows recomputed or moved between blocks,
//                only 2C warm-up rows where a CTA's range of the strip starts), divides, rounds, stores. Lane 0 of the
//                first vertical warp also issues the TMA requests (block n + 2 when block n's fp32 rows are complete:
//                the horizontal warps are then done with that input buffer, no further barrier needed).
// The two passes run concurrently on the same SM: whichever has warps ready keeps the FMA pipe busy.
//
// Division: interior outputs divide by the full kernel sum b, which make_gaussian_kernel leaves within a few ulp of
// 1.0. For such b, a / b == RN (a + a * e) with e ~ 1/b - 1 - ONE FFMA2 - for every a the blur can produce; the
// (b, e) pairs are a whitelist (kOneFmaDiv) verified exhaustively by tests/test_gaussblur_gpu.py through
// b200vf_gauss_selftest_div1. Truncated windows at the frame edges use div2 (reciprocal + two corrections).
//
// Byte-shifted frames (AYUV ...) are blurred as ALIGNED words as in gaussblur_kernel<.., 0> (p0v, patch_w); the two
// pixel columns that straddle the frame edge there (pixel 0, pixel w-1) are written by gauss_lastcol_kernel (gaussblur.cu).
//
// What the ncu captures taught (profiles/r02_gaussblur_stream.md):
//  * every role's hot loop is ~30 KB of straight-line code; the SM's instruction caches hold two such streams, not
//    three or four. Hence ONE copy per role at a time: the edge variant of a pass is chosen per strip / per block for
//    all its warps, never per warp (per-warp variants made the edge strips 1.4-1.5x slower for every warp on the SM);
//  * a warp shuffle inside the unrolled vertical loop makes ptxas guard each step against divergence (BRA.DIV +
//    WARPSYNC) and keep the taps in vector registers; without it the taps live in uniform registers (FMUL2 R, R, UR)
//    and the products of a sample get their own registers. The warp index comes from a shuffle for the same reason;
//  * CTAs get equal WEIGHT, not equal unit counts: see StreamSched.
#pragma once

constexpr int SSTRIP = 128;                 // strip width in aligned columns
constexpr int SBLK = 32;                    // rows per block = slots of the vertical ring
constexpr int SSEG = 32;                    // outputs per horizontal segment
constexpr int SHW = 8, SVW = 8;             // horizontal / vertical warps
constexpr int STMP_NBUF = 2;                // fp32 blocks in flight between the passes (1 = the passes alternate: measured the same)
constexpr int STHREADS = (SHW + SVW) * 32;         // 512: up to 128 registers per thread
constexpr int STMP_PITCH = SSTRIP * 16 + 16; // bytes per fp32 row: +16 so that 16 rows x 16 B cover 256 B of distinct banks
// TMA boxes start at a multiple of 4 pixels (16 bytes): the box of a strip at x0 (a multiple of 4) starts sraw_off (C)
// samples before x0 - C, and the horizontal pass skips them (a compile-time offset in its unrolled stream).
__host__ __device__ constexpr int sraw_off (int C) { return (4 - C % 4) % 4; }
__host__ __device__ constexpr int sraw_w (int C) {          // samples per box row: multiple of 4, /4 odd (LDS.128 of 16 rows conflict-free)
  int w = (SSTRIP + 2 * C + sraw_off (C) + 3) / 4 * 4;
  return ((w / 4) & 1) ? w : w + 4;
}

__device__ __forceinline__ f32x2 mul2v (f32x2 a, f32x2 b) { f32x2 r; asm volatile ("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ void acc2v (f32x2 &acc, f32x2 m, f32x2 one) { asm volatile ("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(m), "l"(one)); }


// The schedule every role walks in the same order: the CTA's contiguous range of (frame, strip, 32-row step) units
// (strip-major, step fastest), cut into segments = runs of consecutive steps of one strip. Ranges have equal WEIGHT:
// strips at the frame's left / right edge cost more per unit (truncated divisors in the horizontal pass: div2 per output
// instead of one FMA), a last strip that is mostly outside the frame costs less; StreamSched describes the strips of a
// frame as four zones of (count, weight in eighths of an interior unit).
struct StreamSched { int zone_strips[4], zone_w[4]; long long frame_weight, total_weight; };
__device__ __forceinline__ int stream_unit_at (const GaussParams &p, const StreamSched &ss, long long t) {
  const int frame = (int) (t / ss.frame_weight);
  long long r = t - (long long) frame * ss.frame_weight;
  int u = frame * p.tiles_x * p.nsteps;
#pragma unroll
  for (int z = 0; z < 4; z++) {
    const long long zw = (long long) ss.zone_strips[z] * p.nsteps * ss.zone_w[z];
    if (r < zw) return u + (int) ((r + ss.zone_w[z] - 1) / ss.zone_w[z]);
    r -= zw; u += ss.zone_strips[z] * p.nsteps;
  }
  return min (u, (frame + 1) * p.tiles_x * p.nsteps);
}
struct SegIter {
  int u, u1;
  int frame, strip, ya, yb, nblocks;
  __device__ __forceinline__ void start (const GaussParams &p, const StreamSched &ss, int block, int nb) {
    u = stream_unit_at (p, ss, ss.total_weight * block / nb);
    u1 = block + 1 == nb ? p.total_units : stream_unit_at (p, ss, ss.total_weight * (block + 1) / nb);
  }
  __device__ __forceinline__ bool next (const GaussParams &p, int C) {
    if (u >= u1) return false;
    const int step = u % p.nsteps;
    strip = (u / p.nsteps) % p.tiles_x;
    frame = u / p.nsteps / p.tiles_x;
    const int nst = min (p.nsteps - step, u1 - u);
    ya = p.y_begin + step * SBLK;
    yb = min (p.y_end, ya + nst * SBLK);
    nblocks = (yb - ya + 2 * C + SBLK - 1) / SBLK;
    u += nst;
    return true;
  }
};

struct StreamConsts {
  f32x2 e2;                 // (e, e): a / b_full == RN (a + a * e)
  f32x2 k2[16];             // taps 0 .. C as (k, k); padded windows carry zero taps first
  StreamSched sched;
};

// ---- horizontal pass of one segment: 32 outputs of row r, channel pair `pair`, from 32 + 2C samples ----
template <int C, bool EDGE>
__device__ __forceinline__ void stream_h_segment (const uint32_t *rawrow, const int sh, uint8_t *tmp_out, const f32x2 (&K)[C + 1],
    const f32x2 e2, const f32x2 one, const float4 *s_divp, const int m0, const int w, const int c_true, const int p0v, const int pair,
    const int i0s, const int iws, const uint32_t wword, const bool use_wword)
{
  constexpr int L = SSEG, NS = L + 2 * C;
  f32x2 A[L];
  uint4 wv = make_uint4 (0, 0, 0, 0);
#pragma unroll
  for (int i = 0; i < NS; i++) {
    constexpr int OFF = sraw_off (C);
    const int iw = i + OFF;                                    // raw word of sample i
    if ((iw & 3) == 0 || i == 0) wv = reinterpret_cast<const uint4 *> (rawrow)[iw / 4];
    uint32_t word = (iw & 3) == 0 ? wv.x : (iw & 3) == 1 ? wv.y : (iw & 3) == 2 ? wv.z : wv.w;
    if (EDGE) {
      // Aligned view of a byte-shifted image ("zero samples outside the frame", see gaussblur_kernel): the low p0v bytes of
      // aligned column 0 are pixel -1 (in memory: the tail of the previous row) and read as zero; of aligned column w only
      // the low p0v bytes exist (pixel w-1), and with unpadded rows they are the first word of the next row (wword).
      const uint32_t keep_hi = 0xffffffffu << (8 * p0v);
      if (i == i0s) word &= keep_hi;
      if (i == iws) word = (use_wword ? wword : word) & ~keep_hi;
    }
    const uint32_t t = word >> sh;
    const f32x2 v = pack2 ((float) (t & 0xffu), (float) ((t >> 8) & 0xffu));
    f32x2 m[C + 1];
#pragma unroll
    for (int j = 0; j <= C; j++) {
      const bool need = (i - j >= 0 && i - j < L) || (i - (2 * C - j) >= 0 && i - (2 * C - j) < L);
      m[j] = need ? mul2v (v, K[j]) : 0ull;
    }
#pragma unroll
    for (int k = 0; k <= 2 * C; k++) {
      const int o = i - k;
      if (o >= 0 && o < L) {
        const int j = k <= C ? k : 2 * C - k;
        if (k == 0) A[o] = m[j]; else acc2v (A[o], m[j], one);
      }
    }
    if (i >= 2 * C) {
      const int o = i - 2 * C;
      f32x2 q;
      if (!EDGE) q = fma2 (A[o], e2, A[o]);
      else {
        // Segment at the frame's left / right edge: only the columns whose window is truncated (or that hold the low
        // p0v bytes of such a pixel) need their own divisors; s_divp holds them per (position, pair) as
        // (-d0, -d1, 1/d0, 1/d1) for this thread's two bytes.
        // (straight-line: every output of an edge segment divides through the table, columns with the full window
        // through its last entry - branches here would cut the unrolled stream into 32 pieces the scheduler cannot overlap)
        const int m_col = m0 + o;
        const int j = m_col <= c_true ? max (m_col, 0) : m_col <= w - 1 - c_true ? 2 * c_true + 2 : min (m_col - (w - 1 - 2 * c_true), 2 * c_true + 1);
        const float4 t = s_divp[2 * j + pair];
        q = div2 (A[o], pack2 (t.x, t.y), pack2 (t.z, t.w));
      }
      *reinterpret_cast<f32x2 *> (tmp_out + o * 16) = q;
    }
  }
}

// ---- vertical pass of one block: 32 samples (rows) of this thread's column / channel pair ----
template <int C, bool EDGE>
__device__ __forceinline__ void stream_v_block (f32x2 (&A)[SBLK], const uint8_t *tmp_in, const f32x2 (&K)[C + 1], const f32x2 e2, const f32x2 one,
    const float2 *s_div, const int o_base, const int ya, const int yb, const int full_h, const int c_true,
    uint8_t *pj /* this thread's two bytes in output row o_base */, const int stride, const bool store)
{
  const f32x2 H2 = 0x3F0000003F000000ull, M2 = 0x4B0000004B000000ull;       // (0.5, 0.5), (2^23, 2^23)
  const unsigned nrows = (unsigned) (yb - ya);
  asm volatile ("" : "+l"(pj));
#pragma unroll
  for (int I = 0; I < SBLK; I++) {
    const f32x2 v = *reinterpret_cast<const f32x2 *> (tmp_in + I * STMP_PITCH);
    f32x2 m[C + 1];
#pragma unroll
    for (int j = 0; j <= C; j++) m[j] = mul2v (v, K[j]);
    A[(I + C) & (SBLK - 1)] = m[0];
#pragma unroll
    for (int k = 1; k <= C; k++) acc2v (A[(I + C - k) & (SBLK - 1)], m[k], one);
#pragma unroll
    for (int k = C - 1; k >= 0; k--) acc2v (A[(I - C + k + SBLK) & (SBLK - 1)], m[k], one);
    // output row o = o_base + I is complete
    const int o = o_base + I;
    const f32x2 a = A[(I - C + SBLK) & (SBLK - 1)];
    f32x2 q;
    if (!EDGE) q = fma2 (a, e2, a);
    else {
      const float2 dv = s_div[edge_index (o, full_h, c_true)];
      q = div2 (a, pack2 (-dv.x, -dv.x), pack2 (dv.y, dv.y));
    }
    const f32x2 y = add2_rm (add2_rm (q, H2), M2);             // finish_word_fast, one channel pair
    const uint32_t two = PRMT ((uint32_t) y, (uint32_t) (y >> 32), 0x0040);      // this thread's two bytes of the word
    // (no shuffle to assemble the 32-bit word in one lane: a shuffle inside this loop makes the compiler guard every
    // step against divergence and keep the taps out of the uniform registers; 16-bit stores merge in L2)
    const unsigned row_ok = (unsigned) (o - ya) < nrows;
    st_u16_if<0> (pj, two, row_ok && store);
    pj += stride;
  }
}

template <int C>
__global__ void __launch_bounds__ (STHREADS, 1)
gaussblur_stream_kernel (const __grid_constant__ CUtensorMap src_map, const __grid_constant__ GaussParams p,
    const __grid_constant__ GaussTaps taps, const __grid_constant__ StreamConsts sc)
{
  constexpr int RAWW = sraw_w (C);
  constexpr int RAW_WORDS = SBLK * RAWW;                    // 32 rows; RAWW * 4 is a multiple of 16, so 32 rows are a multiple of 128 B
  extern __shared__ __align__ (128) float4 smem4[];
  uint32_t *raw = reinterpret_cast<uint32_t *> (smem4);                                     // [2][32][RAWW] u8x4 samples (TMA)
  uint8_t *tmp = reinterpret_cast<uint8_t *> (raw + 2 * RAW_WORDS);                         // [2][32][STMP_PITCH] fp32 rows
  float2 *s_div = reinterpret_cast<float2 *> (tmp + 2 * SBLK * STMP_PITCH);                 // [2c + 1] (sum, 1 / sum), true window
  __shared__ __align__ (8) uint64_t raw_full[2], tmp_full[2], tmp_empty[2];
  const int c = p.center, ws = p.ws;                        // the TRUE window (divisors); C >= c is the padded one
  float4 *s_divp = reinterpret_cast<float4 *> (s_div + 32);                                  // [2c + 2][2 pairs], see stream_h_segment
  for (int i = threadIdx.x; i < ws; i += STHREADS) {
    const float sm = partial_sum (taps.ksum, i, ws, ws, c);
    s_div[i] = make_float2 (sm, __frcp_rn (sm));
  }
  for (int i = threadIdx.x; i < 2 * (ws + 2); i += STHREADS) {
    // position j of a line of exactly ws samples (j = ws: one past the end = aligned column w), channel pair pr:
    // byte b of the column belongs to the pixel at j (b >= p0v) or at j - 1 (b < p0v)
    const int j = i >> 1, pr = i & 1;
    // (entry ws + 1: a column whose two pixels both have the full window)
    const float dv = partial_sum (taps.ksum, j > ws ? c : min (j, ws - 1), ws, ws, c), dl = partial_sum (taps.ksum, j > ws ? c : max (j - 1, 0), ws, ws, c);
    const float d0 = (2 * pr < p.p0v) ? dl : dv, d1 = (2 * pr + 1 < p.p0v) ? dl : dv;
    s_divp[i] = make_float4 (-d0, -d1, __frcp_rn (d0), __frcp_rn (d1));
  }
  if (threadIdx.x == 0) {
    for (int b = 0; b < 2; b++) { mbar_init (&raw_full[b], 1); mbar_init (&tmp_full[b], SHW); mbar_init (&tmp_empty[b], SVW); }
    mbar_fence_init ();
  }
  __syncthreads ();

  // warp index through a shuffle: the compiler then knows the role branches below are warp-uniform (uniform registers
  // for the taps, no divergence guards around the shuffles and barriers inside them)
  const int warp = __shfl_sync (0xffffffffu, (int) (threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  f32x2 K[C + 1];
#pragma unroll
  for (int j = 0; j <= C; j++) K[j] = sc.k2[j];
  const f32x2 e2 = sc.e2, one = p.one2;
  SegIter it;
  it.start (p, sc.sched, blockIdx.x, gridDim.x);
  const int p0v = p.p0v;
  int n = 0;                                                // running block number; its buffers are n & 1

  if (warp < SHW) {
    // ---------------------------------------------------------------- horizontal pass
    // segment g of row half (warp >> 2); warps w and w + 4 share a scheduler (w & 3), so the second half is rotated by one:
    // the segment at a frame edge (heavier: truncated divisors) then lands on two different schedulers
    const int g = (warp + (warp >> 2)) & 3, r = (warp >> 2) * 16 + (lane >> 1), pair = lane & 1;
    const int sh = 16 * pair;
    while (it.next (p, C)) {
      const int tx0 = p.x_tile0 + it.strip * SSTRIP;
      const int m0 = tx0 + g * SSEG;                        // aligned column of this segment's output 0
      // no column of the segment is closer than c to a frame edge (the low p0v bytes of a column belong to the pixel on its left)
      // no column of the STRIP is closer than c to a frame edge (the low p0v bytes of a column belong to the pixel on its
      // left). Decided per strip, not per segment: all eight warps then run the same copy of the unrolled stream, and with
      // the vertical pass's at most two instruction streams are live on the SM (the copies are ~30 KB each; a third and
      // fourth one executing concurrently in the edge strips cost more in instruction-cache misses than the table-driven
      // division costs the segments that would not have needed it)
      const bool interior = tx0 - (p0v ? 1 : 0) >= c && tx0 + SSTRIP <= p.w - c;
      const bool outside = m0 >= p.ncols || m0 + SSEG <= 0 || m0 >= p.x_end || m0 + SSEG <= p.x_begin;
      // aligned view: raw columns of aligned columns 0 and w, if this strip's box holds them (see gaussblur_kernel)
      // sample numbers (within this segment's stream) of aligned columns 0 and w, when the image is byte-shifted and the
      // stream holds them; only edge segments can (both lie within C columns of an output column next to the frame edge)
      const int i0s = p0v && C - m0 >= 0 && C - m0 < SSEG + 2 * C ? C - m0 : -1;
      const int iws = p0v && p.w + C - m0 >= 0 && p.w + C - m0 < SSEG + 2 * C ? p.w + C - m0 : -1;
      const uint8_t *srcf = p.src + (size_t) it.frame * p.src_frame_stride;
      for (int j = 0; j < it.nblocks; j++, n++) {
        const int b = n & 1, k = n >> 1;
        mbar_wait (&raw_full[b], k & 1);
        uint32_t *rawb = raw + b * RAW_WORDS;
        // column w of an unpadded byte-shifted frame = first word of the next physical row (bytes past the readable range
        // read as 0): fetched now, used near the end of the segment's stream
        uint32_t wword = 0;
        if (iws >= 0 && p.patch_w) {
          const int gr = it.ya - C + j * SBLK + r;            // global row of this thread's raw row
          const long long a = (long long) (gr + 1 - p.row0) * p.stride;
          if (gr >= 0 && gr < p.full_h && a >= p.in_lo && a + 4 <= p.in_hi) wword = ldg_u32 (srcf + a);
        }
        const int tb = STMP_NBUF == 2 ? b : 0, tk = STMP_NBUF == 2 ? k : n;
        if (tk > 0) mbar_wait (&tmp_empty[tb], (tk - 1) & 1);
        const uint32_t *rawrow = rawb + r * RAWW + g * SSEG;
        uint8_t *tmp_out = tmp + (size_t) tb * SBLK * STMP_PITCH + r * STMP_PITCH + g * SSEG * 16 + pair * 8;
        if (outside) {}                      // no column of the segment exists: nobody stores what the vertical pass makes of it
        else if (interior) stream_h_segment<C, false> (rawrow, sh, tmp_out, K, e2, one, s_divp, m0, p.w, c, p0v, pair, -1, -1, 0u, false);
        else stream_h_segment<C, true> (rawrow, sh, tmp_out, K, e2, one, s_divp, m0, p.w, c, p0v, pair, i0s, iws, wword, p.patch_w != 0);
        __syncwarp ();
        if (lane == 0) mbar_arrive (&tmp_full[tb]);            // also: this warp is done with raw[b]
      }
    }
    return;
  }

  // ------------------------------------------------------------------ vertical pass
  const int tv = threadIdx.x - SHW * 32;                    // 0 .. 255 = (column, pair)
  const int xi = tv >> 1;
  f32x2 A[SBLK];
#pragma unroll
  for (int i = 0; i < SBLK; i++) A[i] = 0ull;
  // Producer duty (first vertical warp, lane 0): block n + 2 is requested right after tmp_full of block n was seen -
  // all horizontal warps have then finished reading raw[n & 1], so the box can land there without a further barrier,
  // a whole block time before it is needed.
  const bool producer = threadIdx.x == SHW * 32;
  SegIter pit = it;
  int pblk = 0, pvalid = 0;
  auto issue_next = [&] (int nb) {                            // request the next block of the schedule into buffer nb & 1
    if (!pvalid || ++pblk >= pit.nblocks) { pvalid = pit.next (p, C); pblk = 0; }
    if (!pvalid) return;
    const int b = nb & 1;
    mbar_expect_tx (&raw_full[b], RAW_WORDS * 4);
    tma_load_3d (raw + b * RAW_WORDS, &src_map, &raw_full[b], p.x_tile0 + pit.strip * SSTRIP - C - sraw_off (C), pit.ya - C + pblk * SBLK - p.buf_row0, pit.frame);
  };
  if (producer) { issue_next (0); issue_next (1); }
  while (it.next (p, C)) {
    const int tx0 = p.x_tile0 + it.strip * SSTRIP;
    const int xg = tx0 + xi;                                // this thread's aligned column
    uint8_t *dst = p.dst + (size_t) it.frame * p.frame_stride;
    // This thread's two bytes (2 * pair, 2 * pair + 1 of aligned column xg) are stored when the column lies in the launch's
    // region and both bytes belong to a pixel of the frame: byte i is pixel xg's (i >= p0v) or pixel xg - 1's (i < p0v).
    // Columns 0 and w of a byte-shifted image hold bytes of one pixel only: those two pixel columns are written by
    // gauss_lastcol_kernel (gaussblur.cu), not here.
    const int pair = tv & 1;
    bool store = xg >= p.x_begin && xg < p.x_end;
    for (int i = 2 * pair; i < 2 * pair + 2; i++) store = store && (i >= p0v ? (xg >= 0 && xg < p.w) : (xg >= 1 && xg <= p.w));
    const int wx0 = tx0 + (warp - SHW) * 16;
    const bool warp_outside = wx0 >= min (p.x_end, p.ncols) || wx0 + 16 <= max (p.x_begin, 0);
    for (int j = 0; j < it.nblocks; j++, n++) {
      const int b = n & 1, k = n >> 1;
      const int o_base = it.ya - 2 * C + j * SBLK;          // output row completed by the block's first sample
      const bool interior = o_base >= c && o_base + SBLK <= p.full_h - c;
      uint8_t *pj = dst + ((long long) (o_base - p.row0) * p.stride + 4ll * xg + 2 * pair);
      const int tb = STMP_NBUF == 2 ? b : 0, tk = STMP_NBUF == 2 ? k : n;
      mbar_wait (&tmp_full[tb], tk & 1);
      if (producer) issue_next (n + 2);
      const uint8_t *tmp_in = tmp + (size_t) tb * SBLK * STMP_PITCH + tv * 8;
      if (warp_outside) {}                   // none of this warp's 16 columns is stored
      else if (interior) stream_v_block<C, false> (A, tmp_in, K, e2, one, s_div, o_base, it.ya, it.yb, p.full_h, c, pj, p.stride, store);
      else stream_v_block<C, true> (A, tmp_in, K, e2, one, s_div, o_base, it.ya, it.yb, p.full_h, c, pj, p.stride, store);
      __syncwarp ();
      if (lane == 0) mbar_arrive (&tmp_empty[tb]);
    }
  }
}
