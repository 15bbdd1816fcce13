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
//   warps 0-7    horizontal pass: input by TMA, one box (32 rows x (128 + 2C) aligned u8x4 samples, zero-filled outside
//                the frame) per block into a 2-deep ring (issued by the first vertical warp, see below); lane = (row, channel pair), warp = 16 rows of one 32-pixel segment; a thread
//                streams over the 32 + 2C samples of its segment (fully unrolled: products that no output of the
//                segment needs are never formed), divides, and writes fp32 rows into a 2-deep ring of 32-row blocks;
//   warps 8-15   vertical pass: lane = (column, channel pair) of a 128-pixel strip; a thread streams DOWN the strip,
//                its ring lives in registers for the whole walk (no halo rows recomputed or moved between blocks,
//                only 2C warm-up rows where a CTA's range of the strip starts), divides, rounds, stores.
// The two passes run concurrently on the same SM: whichever has warps ready keeps the FMA pipe busy.
//
// Division: interior outputs divide by the full kernel sum b, which make_gaussian_kernel leaves within a few ulp of
// 1.0. For such b, a / b == RN (a + a * e) with e ~ 1/b - 1 - ONE FFMA2 - for every a the blur can produce; the
// (b, e) pairs are a whitelist (kOneFmaDiv) verified exhaustively by tests/test_gaussblur_gpu.py through
// b200vf_gauss_selftest_div1. Truncated windows at the frame edges use div2 (reciprocal + two corrections).
//
// Byte-shifted frames (AYUV ...) are blurred as ALIGNED words exactly as in gaussblur_kernel<.., 0> (p0v, patch_w).
#pragma once

constexpr int SSTRIP = 128;                 // strip width in aligned columns
constexpr int SBLK = 32;                    // rows per block = slots of the vertical ring
constexpr int SSEG = 32;                    // outputs per horizontal segment
constexpr int SHW = 8, SVW = 8;             // horizontal / vertical warps
constexpr int STHREADS = (SHW + SVW) * 32;         // 512: up to 128 registers per thread
constexpr int STMP_PITCH = SSTRIP * 16 + 16; // bytes per fp32 row: +16 so that 16 rows x 16 B cover 256 B of distinct banks
__host__ __device__ constexpr int sraw_w (int C) {          // samples per box row: multiple of 4, /4 odd (LDS.128 of 16 rows conflict-free)
  int w = (SSTRIP + 2 * C + 3) / 4 * 4;
  return ((w / 4) & 1) ? w : w + 4;
}

__device__ __forceinline__ f32x2 mul2v (f32x2 a, f32x2 b) { f32x2 r; asm volatile ("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ void acc2v (f32x2 &acc, f32x2 m, f32x2 one) { asm volatile ("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(m), "l"(one)); }
__device__ __forceinline__ void named_bar_sync (int id, int nthreads) { asm volatile ("bar.sync %0, %1;" :: "r"(id), "r"(nthreads) : "memory"); }

struct StreamConsts {
  f32x2 e2;                 // (e, e): a / b_full == RN (a + a * e)
  f32x2 k2[16];             // taps 0 .. C as (k, k); padded windows carry zero taps first
};

// The schedule every role walks in the same order: the CTA's contiguous range of (frame, strip, 32-row step) units,
// cut into segments = runs of consecutive steps of one strip.
struct SegIter {
  int u, u1, nsteps, tiles_x;
  int frame, strip, ya, yb, nblocks;
  __device__ __forceinline__ bool next (const GaussParams &p, int C) {
    if (u >= u1) return false;
    const int step = u % nsteps;
    strip = (u / nsteps) % tiles_x;
    frame = u / nsteps / tiles_x;
    const int nst = min (nsteps - step, u1 - u);
    ya = p.y_begin + step * SBLK;
    yb = min (p.y_end, ya + nst * SBLK);
    nblocks = (yb - ya + 2 * C + SBLK - 1) / SBLK;
    u += nst;
    return true;
  }
};

// ---- horizontal pass of one segment: 32 outputs of row r, channel pair `pair`, from 32 + 2C samples ----
template <int C>
__device__ __forceinline__ void stream_h_segment (const uint32_t *rawrow, const int sh, uint8_t *tmp_out, const StreamConsts &sc,
    const f32x2 one, const bool interior, const float2 *s_div, const int m0, const int w, const int c_true, const int p0v, const int pair)
{
  constexpr int L = SSEG, NS = L + 2 * C;
  f32x2 A[L];
  uint4 wv = make_uint4 (0, 0, 0, 0);
#pragma unroll
  for (int i = 0; i < NS; i++) {
    if ((i & 3) == 0) wv = reinterpret_cast<const uint4 *> (rawrow)[i / 4];
    const uint32_t word = (i & 3) == 0 ? wv.x : (i & 3) == 1 ? wv.y : (i & 3) == 2 ? wv.z : wv.w;
    const uint32_t t = word >> sh;
    const f32x2 v = pack2 ((float) (t & 0xffu), (float) ((t >> 8) & 0xffu));
    f32x2 m[C + 1];
#pragma unroll
    for (int j = 0; j <= C; j++) {
      const bool need = (i - j >= 0 && i - j < L) || (i - (2 * C - j) >= 0 && i - (2 * C - j) < L);
      m[j] = need ? mul2v (v, sc.k2[j]) : 0ull;
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
      if (interior) q = fma2 (A[o], sc.e2, A[o]);
      else {
        // divisor of the pixel a byte belongs to: pixel m, or m - 1 for the low p0v bytes of an aligned column
        const int m_col = m0 + o;
        const float2 dv = s_div[edge_index (m_col, w, c_true)];
        const float2 dl = p0v ? s_div[edge_index (m_col - 1, w, c_true)] : dv;
        const float2 d0 = (2 * pair < p0v) ? dl : dv, d1 = (2 * pair + 1 < p0v) ? dl : dv;
        q = div2 (A[o], pack2 (-d0.x, -d1.x), pack2 (d0.y, d1.y));
      }
      *reinterpret_cast<f32x2 *> (tmp_out + o * 16) = q;
    }
  }
}

// ---- vertical pass of one block: 32 samples (rows) of this thread's column / channel pair ----
template <int C>
__device__ __forceinline__ void stream_v_block (f32x2 (&A)[SBLK], const uint8_t *tmp_in, const StreamConsts &sc, const f32x2 one,
    const bool interior, const float2 *s_div, const int o_base, const int ya, const int yb, const int full_h, const int c_true,
    uint8_t *pj /* this column's word in output row o_base */, const int stride, const unsigned bm, const bool partial_warp,
    const int skip_row)
{
  const f32x2 H2 = 0x3F0000003F000000ull, M2 = 0x4B0000004B000000ull;       // (0.5, 0.5), (2^23, 2^23)
  const unsigned nrows = (unsigned) (yb - ya);
  asm volatile ("" : "+l"(pj));
#pragma unroll
  for (int I = 0; I < SBLK; I++) {
    const f32x2 v = *reinterpret_cast<const f32x2 *> (tmp_in + I * STMP_PITCH);
    f32x2 m[C + 1];
#pragma unroll
    for (int j = 0; j <= C; j++) m[j] = mul2v (v, sc.k2[j]);
    A[(I + C) & (SBLK - 1)] = m[0];
#pragma unroll
    for (int k = 1; k <= C; k++) acc2v (A[(I + C - k) & (SBLK - 1)], m[k], one);
#pragma unroll
    for (int k = C - 1; k >= 0; k--) acc2v (A[(I - C + k + SBLK) & (SBLK - 1)], m[k], one);
    // output row o = o_base + I is complete
    const int o = o_base + I;
    const f32x2 a = A[(I - C + SBLK) & (SBLK - 1)];
    f32x2 q;
    if (interior) q = fma2 (a, sc.e2, a);
    else {
      const float2 dv = s_div[edge_index (o, full_h, c_true)];
      q = div2 (a, pack2 (-dv.x, -dv.x), pack2 (dv.y, dv.y));
    }
    const f32x2 y = add2_rm (add2_rm (q, H2), M2);             // finish_word_fast, one channel pair
    const uint32_t two = PRMT ((uint32_t) y, (uint32_t) (y >> 32), 0x0040);
    const uint32_t other = __shfl_xor_sync (0xffffffffu, two, 1);
    const uint32_t word = PRMT (two, other, 0x5410);           // even lane: bytes 0-1 its own, 2-3 the odd lane's
    const unsigned row_ok = (unsigned) (o - ya) < nrows;
    if (!partial_warp) st_u32_if<0> (pj, word, row_ok && bm == 0xfu);
    else {
      // a warp that owns aligned column 0 or w of a byte-shifted image (bytes of one pixel only): bytewise there.
      // skip_row: with unpadded rows, column w of the shard's last row is the next row's first word - not ours.
      const unsigned ok = row_ok && !(o == skip_row);
      st_u32_if<0> (pj, word, ok && bm == 0xfu);
      const unsigned part = ok && bm != 0xfu;
      st_u8_if<0> (pj, word, part && (bm & 1u));
      st_u8_if<1> (pj, word >> 8, part && (bm & 2u));
      st_u8_if<2> (pj, word >> 16, part && (bm & 4u));
      st_u8_if<3> (pj, word >> 24, part && (bm & 8u));
    }
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
  for (int i = threadIdx.x; i < ws; i += STHREADS) {
    const float sm = partial_sum (taps.ksum, i, ws, ws, c);
    s_div[i] = make_float2 (sm, __frcp_rn (sm));
  }
  if (threadIdx.x == 0) {
    for (int b = 0; b < 2; b++) { mbar_init (&raw_full[b], 1); mbar_init (&tmp_full[b], SHW); mbar_init (&tmp_empty[b], SVW); }
    mbar_fence_init ();
  }
  __syncthreads ();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  SegIter it;
  it.u = (int) ((long long) p.total_units * blockIdx.x / gridDim.x);
  it.u1 = (int) ((long long) p.total_units * (blockIdx.x + 1) / gridDim.x);
  it.nsteps = p.nsteps; it.tiles_x = p.tiles_x;
  const int p0v = p.p0v;
  int n = 0;                                                // running block number; its buffers are n & 1

  if (warp < SHW) {
    // ---------------------------------------------------------------- horizontal pass
    const int g = warp & 3, r = (warp >> 2) * 16 + (lane >> 1), pair = lane & 1;
    const int sh = 16 * pair;
    while (it.next (p, C)) {
      const int tx0 = p.x_tile0 + it.strip * SSTRIP;
      const int m0 = tx0 + g * SSEG;                        // aligned column of this segment's output 0
      // no column of the segment is closer than c to a frame edge (the low p0v bytes of a column belong to the pixel on its left)
      const bool interior = m0 - (p0v ? 1 : 0) >= c && m0 + SSEG <= p.w - c;
      // aligned view: raw columns of aligned columns 0 and w, if this strip's box holds them (see gaussblur_kernel)
      const int i0 = C - tx0, iw = p.w + C - tx0;
      const bool has0 = p0v && i0 >= 0 && i0 < RAWW, hasw = p0v && iw >= 0 && iw < RAWW;
      const uint8_t *srcf = p.src + (size_t) it.frame * p.src_frame_stride;
      for (int j = 0; j < it.nblocks; j++, n++) {
        const int b = n & 1, k = n >> 1;
        mbar_wait (&raw_full[b], k & 1);
        uint32_t *rawb = raw + b * RAW_WORDS;
        if (has0 || hasw) {                                 // uniform over the 8 warps
          if (g == 0 && pair == 0) {
            const uint32_t keep_hi = 0xffffffffu << (8 * p0v);
            uint32_t *rowp = rawb + r * RAWW;
            if (has0) rowp[i0] &= keep_hi;
            if (hasw) {
              uint32_t v = rowp[iw];
              if (p.patch_w) {
                const int gr = it.ya - C + j * SBLK + r;    // global row of this raw row
                const long long a = (long long) (gr + 1 - p.row0) * p.stride;
                v = (gr >= 0 && gr < p.full_h && a >= p.in_lo && a + 4 <= p.in_hi) ? ldg_u32 (srcf + a) : 0u;
              }
              rowp[iw] = v & ~keep_hi;
            }
          }
          named_bar_sync (1, SHW * 32);
        }
        if (k > 0) mbar_wait (&tmp_empty[b], (k - 1) & 1);
        stream_h_segment<C> (rawb + r * RAWW + g * SSEG, sh, tmp + (size_t) b * SBLK * STMP_PITCH + r * STMP_PITCH + g * SSEG * 16 + pair * 8,
            sc, p.one2, interior, s_div, m0, p.w, c, p0v, pair);
        __syncwarp ();
        if (lane == 0) mbar_arrive (&tmp_full[b]);             // also: this warp is done with raw[b]
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
    if (!pvalid) { pit.u = pit.u1; return; }
    const int b = nb & 1;
    mbar_expect_tx (&raw_full[b], RAW_WORDS * 4);
    tma_load_3d (raw + b * RAW_WORDS, &src_map, &raw_full[b], p.x_tile0 + pit.strip * SSTRIP - C, pit.ya - C + pblk * SBLK - p.buf_row0, pit.frame);
  };
  if (producer) { issue_next (0); issue_next (1); }
  while (it.next (p, C)) {
    const int tx0 = p.x_tile0 + it.strip * SSTRIP;
    const int xg = tx0 + xi;                                // this thread's aligned column
    uint8_t *dst = p.dst + (size_t) it.frame * p.frame_stride;
    // byte i of aligned column xg belongs to pixel xg (i >= p0v) or xg - 1 (i < p0v): stored (by the even lane of the
    // pair, which assembles the word) if that pixel exists and the column lies in this launch's region
    unsigned bm = 0;
    if (!(lane & 1) && xg >= p.x_begin && xg < p.x_end)
      for (int i = 0; i < 4; i++)
        if (i >= p0v ? (xg >= 0 && xg < p.w) : (xg >= 1 && xg <= p.w)) bm |= 1u << i;
    const bool partial_warp = __any_sync (0xffffffffu, bm != 0 && bm != 0xfu);
    // unpadded rows: column w of row r is the first word of row r + 1; on the shard's last row that word is not ours
    const int skip_row = (p.patch_w && xg == p.w) ? p.row0 + (int) (p.out_hi / p.stride) - 1 : -0x40000000;
    for (int j = 0; j < it.nblocks; j++, n++) {
      const int b = n & 1, k = n >> 1;
      const int o_base = it.ya - 2 * C + j * SBLK;          // output row completed by the block's first sample
      const bool interior = o_base >= c && o_base + SBLK <= p.full_h - c;
      uint8_t *pj = dst + ((long long) (o_base - p.row0) * p.stride + 4ll * xg);
      mbar_wait (&tmp_full[b], k & 1);
      if (producer) issue_next (n + 2);
      stream_v_block<C> (A, tmp + (size_t) b * SBLK * STMP_PITCH + tv * 8, sc, p.one2, interior, s_div, o_base, it.ya, it.yb, p.full_h, c,
          pj, p.stride, bm, partial_warp, skip_row);
      __syncwarp ();
      if (lane == 0) mbar_arrive (&tmp_empty[b]);
    }
  }
}
