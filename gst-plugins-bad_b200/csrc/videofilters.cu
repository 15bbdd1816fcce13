// videofilters.cu - the `videofiltersbad` plugin's per-pixel loops for sm_100a (the sibling
// filters with the same GstVideoFilter boundary, on planar / packed YUV instead of packed RGB).
//
//   zebrastripe  gst_zebra_stripe_transform_frame_ip     gst/videofilters/gstzebrastripe.c:205-253
//   videodiff    gst_video_diff_transform_frame_ip_planarY gst/videofilters/gstvideodiff.c:94-129 (luma plane; the
//                chroma planes are plain copies)
//   scenechange  get_frame_score -> orc_sad_nxm_u8       gst/videofilters/gstscenechange.c:141-155,
//                gstscenechangeorc.orc (`accsadubl`), C backup gstscenechangeorc-dist.c:147-180;
//                the decision that follows (:196-236) is host arithmetic on five doubles (b200vf_scenechange_update).
//
// All three are byte streams bound by HBM: 2 (zebrastripe, in place), 3 (videodiff) and 2 (SAD, read only) bytes of
// traffic per luma sample. A thread owns the same 16-byte column of four consecutive rows, so four independent
// 128-bit loads per source are in flight before the first byte is used; the per-byte tests are done four bytes per
// instruction with the SIMD-in-word video instructions (vcmpgeu4 / vabsdiffu4 / vsadu4). Rows whose base or pitch
// is not 16-byte aligned take the same code on 32-bit words; the last width % 4 samples of a row are done bytewise.
#include "common.cuh"
#include <stdlib.h>

namespace {

constexpr int VF_ROWS = 4;                  // rows per thread
constexpr int VF_TX = 64, VF_TY = 4;        // threads per CTA: 64 columns x 4 row groups

// 0xff in every byte lane whose pixel index i (lane b of the word holds pixel i0 + b) has (i + c) & 4 set
__device__ __forceinline__ uint32_t stripe_mask4 (int i0, int c) {
  const uint32_t s = (uint32_t) (i0 + c);
  const uint32_t base = (s & 4u) ? 0xffffffffu : 0u;
  const uint32_t k = s & 3u;                                 // lanes b >= 4 - k carry into bit 2
  const uint32_t flip = k ? (0xffffffffu << (8u * (4u - k))) : 0u;
  return base ^ flip;
}

// ---- zebrastripe --------------------------------------------------------------------------------------
// planar luma (pixel_stride 1): word = 4 consecutive pixels
__device__ __forceinline__ uint32_t zebra_word (uint32_t v, uint32_t thr4, int i0, int c) {
  const uint32_t m = __vcmpgeu4 (v, thr4) & stripe_mask4 (i0, c);
  return (v & ~m) | (0x10101010u & m);
}
__device__ __forceinline__ uint8_t zebra_byte (uint8_t v, int thr, int i, int c) {
  return (v >= thr && ((i + c) & 4)) ? (uint8_t) 16 : v;
}

template <int WORDS>   // words per thread per row: 4 (128-bit accesses) or 1
__global__ void __launch_bounds__ (VF_TX * VF_TY)
zebra_planar_kernel (uint8_t *luma, int row_stride, size_t frame_stride, int width, int height, int thr, int t)
{
  const int w0 = (blockIdx.x * VF_TX + threadIdx.x) * WORDS;           // first word of this thread's column
  const int j0 = (blockIdx.y * VF_TY + threadIdx.y) * VF_ROWS;
  const int nwords = width >> 2;
  uint8_t *base = luma + (size_t) blockIdx.z * frame_stride;
  const int tt = t + blockIdx.z;                                       // the element counts t up once per frame (:217)
  const uint32_t thr4 = (uint32_t) min (thr, 255) * 0x01010101u;
  const bool none = thr > 255;                                         // y_threshold <= 235 for threshold <= 100; kept for any input
  if (w0 < nwords && !none) {
    uint32_t v[VF_ROWS][WORDS];
#pragma unroll
    for (int r = 0; r < VF_ROWS; r++) {
      if (j0 + r >= height) break;
      const uint32_t *p = reinterpret_cast<const uint32_t *> (base + (size_t) (j0 + r) * row_stride) + w0;
      if (WORDS == 4 && w0 + 4 <= nwords) { const uint4 q = *reinterpret_cast<const uint4 *> (p); v[r][0] = q.x; v[r][1] = q.y; v[r][2] = q.z; v[r][3] = q.w; }
      else
#pragma unroll
        for (int k = 0; k < WORDS; k++) if (w0 + k < nwords) v[r][k] = p[k];
    }
#pragma unroll
    for (int r = 0; r < VF_ROWS; r++) {
      if (j0 + r >= height) break;
      uint32_t *p = reinterpret_cast<uint32_t *> (base + (size_t) (j0 + r) * row_stride) + w0;
      uint32_t o[WORDS];
#pragma unroll
      for (int k = 0; k < WORDS; k++) o[k] = zebra_word (v[r][k], thr4, 4 * (w0 + k), j0 + r + tt);
      if (WORDS == 4 && w0 + 4 <= nwords) *reinterpret_cast<uint4 *> (p) = make_uint4 (o[0], o[1], o[2], o[3]);
      else
#pragma unroll
        for (int k = 0; k < WORDS; k++) if (w0 + k < nwords) p[k] = o[k];
    }
  }
  // the last width % 4 samples of each row
  if (blockIdx.x == 0 && threadIdx.x < (width & 3) && !none) {
    const int i = (nwords << 2) + threadIdx.x;
    for (int r = 0; r < VF_ROWS; r++) {
      if (j0 + r >= height) break;
      uint8_t *p = base + (size_t) (j0 + r) * row_stride + i;
      *p = zebra_byte (*p, thr, i, j0 + r + tt);
    }
  }
}

// packed formats: the luma sample of pixel i is byte i * ps of the row that starts at `luma` (ps = 2: YUY2 / UYVY,
// ps = 4: AYUV). One pixel per thread and row; the other bytes are not touched (bytewise read-modify-write).
__global__ void __launch_bounds__ (VF_TX * VF_TY)
zebra_packed_kernel (uint8_t *luma, int ps, int row_stride, size_t frame_stride, int width, int height, int thr, int t)
{
  const int i = blockIdx.x * VF_TX + threadIdx.x;
  const int j0 = (blockIdx.y * VF_TY + threadIdx.y) * VF_ROWS;
  if (i >= width) return;
  uint8_t *base = luma + (size_t) blockIdx.z * frame_stride + (size_t) i * ps;
  const int tt = t + blockIdx.z;
  uint8_t v[VF_ROWS];
#pragma unroll
  for (int r = 0; r < VF_ROWS; r++) if (j0 + r < height) v[r] = base[(size_t) (j0 + r) * row_stride];
#pragma unroll
  for (int r = 0; r < VF_ROWS; r++)
    if (j0 + r < height && v[r] >= thr && ((i + j0 + r + tt) & 4)) base[(size_t) (j0 + r) * row_stride] = 16;
}

// ---- videodiff ----------------------------------------------------------------------------------------
// (s2 < s1 - threshold) || (s2 > s1 + threshold) in int arithmetic == |s2 - s1| > threshold
__device__ __forceinline__ uint32_t diff_word (uint32_t s1, uint32_t s2, uint32_t thr4, bool none, int i0, int c) {
  const uint32_t m = none ? 0u : __vcmpgtu4 (__vabsdiffu4 (s1, s2), thr4);
  const uint32_t mark = (0x10101010u & stripe_mask4 (i0, c)) | (0xf0f0f0f0u & ~stripe_mask4 (i0, c));     // 16 on the stripe, else 240
  return (s2 & ~m) | (mark & m);
}

template <int WORDS>
__global__ void __launch_bounds__ (VF_TX * VF_TY)
videodiff_kernel (const uint8_t *s_old, int old_stride, size_t old_fs, const uint8_t *s_new, int new_stride, size_t new_fs,
    uint8_t *dst, int dst_stride, size_t dst_fs, int width, int height, int threshold, int t)
{
  const int w0 = (blockIdx.x * VF_TX + threadIdx.x) * WORDS;
  const int j0 = (blockIdx.y * VF_TY + threadIdx.y) * VF_ROWS;
  const int nwords = width >> 2;
  const uint8_t *a = s_old + (size_t) blockIdx.z * old_fs, *b = s_new + (size_t) blockIdx.z * new_fs;
  uint8_t *d = dst + (size_t) blockIdx.z * dst_fs;
  const bool none = threshold >= 255;                       // no byte difference exceeds 255
  const uint32_t thr4 = (uint32_t) max (0, min (threshold, 255)) * 0x01010101u;
  const bool all = threshold < 0;                           // every sample differs by more than a negative threshold
  if (w0 < nwords) {
    uint32_t va[VF_ROWS][WORDS], vb[VF_ROWS][WORDS];
#pragma unroll
    for (int r = 0; r < VF_ROWS; r++) {
      if (j0 + r >= height) break;
      const uint32_t *pa = reinterpret_cast<const uint32_t *> (a + (size_t) (j0 + r) * old_stride) + w0;
      const uint32_t *pb = reinterpret_cast<const uint32_t *> (b + (size_t) (j0 + r) * new_stride) + w0;
      if (WORDS == 4 && w0 + 4 <= nwords) {
        const uint4 qa = ld_stream_v4 (pa), qb = ld_stream_v4 (pb);
        va[r][0] = qa.x; va[r][1] = qa.y; va[r][2] = qa.z; va[r][3] = qa.w;
        vb[r][0] = qb.x; vb[r][1] = qb.y; vb[r][2] = qb.z; vb[r][3] = qb.w;
      } else
#pragma unroll
        for (int k = 0; k < WORDS; k++) if (w0 + k < nwords) { va[r][k] = ldg_u32 (pa + k); vb[r][k] = ldg_u32 (pb + k); }
    }
#pragma unroll
    for (int r = 0; r < VF_ROWS; r++) {
      if (j0 + r >= height) break;
      uint32_t *pd = reinterpret_cast<uint32_t *> (d + (size_t) (j0 + r) * dst_stride) + w0;
      uint32_t o[WORDS];
#pragma unroll
      for (int k = 0; k < WORDS; k++) {
        const int i0 = 4 * (w0 + k), c = j0 + r + t;
        o[k] = all ? ((0x10101010u & stripe_mask4 (i0, c)) | (0xf0f0f0f0u & ~stripe_mask4 (i0, c)))
                   : diff_word (va[r][k], vb[r][k], thr4, none, i0, c);
      }
      if (WORDS == 4 && w0 + 4 <= nwords) st_stream_v4 (pd, make_uint4 (o[0], o[1], o[2], o[3]));
      else
#pragma unroll
        for (int k = 0; k < WORDS; k++) if (w0 + k < nwords) pd[k] = o[k];
    }
  }
  if (blockIdx.x == 0 && threadIdx.x < (width & 3)) {
    const int i = (nwords << 2) + threadIdx.x;
    for (int r = 0; r < VF_ROWS; r++) {
      if (j0 + r >= height) break;
      const int s1 = a[(size_t) (j0 + r) * old_stride + i], s2 = b[(size_t) (j0 + r) * new_stride + i];
      uint8_t o = (uint8_t) s2;
      if ((s2 < s1 - threshold) || (s2 > s1 + threshold)) o = ((i + j0 + r + t) & 4) ? 16 : 240;
      d[(size_t) (j0 + r) * dst_stride + i] = o;
    }
  }
}

// ---- sum of absolute differences ----------------------------------------------------------------------
// sums[frame] += SAD of this thread's samples; the accumulator is the reference's orc_uint32: it wraps modulo 2^32
// (an 8K frame can exceed it), and so does this one (32-bit partial sums, 32-bit atomicAdd).
template <int WORDS>
__global__ void __launch_bounds__ (VF_TX * VF_TY)
sad_kernel (const uint8_t *a, int a_stride, size_t a_fs, const uint8_t *b, int b_stride, size_t b_fs, int width, int height,
    uint32_t *sums)
{
  const int w0 = (blockIdx.x * VF_TX + threadIdx.x) * WORDS;
  const int j0 = (blockIdx.y * VF_TY + threadIdx.y) * VF_ROWS;
  const int nwords = width >> 2;
  const uint8_t *fa = a + (size_t) blockIdx.z * a_fs, *fb = b + (size_t) blockIdx.z * b_fs;
  uint32_t acc = 0;
  if (w0 < nwords) {
    uint32_t va[VF_ROWS][WORDS], vb[VF_ROWS][WORDS];
#pragma unroll
    for (int r = 0; r < VF_ROWS; r++) {
#pragma unroll
      for (int k = 0; k < WORDS; k++) { va[r][k] = 0; vb[r][k] = 0; }
      if (j0 + r >= height) continue;
      const uint32_t *pa = reinterpret_cast<const uint32_t *> (fa + (size_t) (j0 + r) * a_stride) + w0;
      const uint32_t *pb = reinterpret_cast<const uint32_t *> (fb + (size_t) (j0 + r) * b_stride) + w0;
      if (WORDS == 4 && w0 + 4 <= nwords) {
        const uint4 qa = ld_stream_v4 (pa), qb = ld_stream_v4 (pb);
        va[r][0] = qa.x; va[r][1] = qa.y; va[r][2] = qa.z; va[r][3] = qa.w;
        vb[r][0] = qb.x; vb[r][1] = qb.y; vb[r][2] = qb.z; vb[r][3] = qb.w;
      } else
#pragma unroll
        for (int k = 0; k < WORDS; k++) if (w0 + k < nwords) { va[r][k] = ldg_u32 (pa + k); vb[r][k] = ldg_u32 (pb + k); }
    }
#pragma unroll
    for (int r = 0; r < VF_ROWS; r++)
#pragma unroll
      for (int k = 0; k < WORDS; k++) acc = __vsadu4 (va[r][k], vb[r][k]) + acc;
  }
  if (blockIdx.x == 0 && threadIdx.x < (width & 3)) {
    const int i = (nwords << 2) + threadIdx.x;
    for (int r = 0; r < VF_ROWS; r++) {
      if (j0 + r >= height) break;
      acc += (uint32_t) abs ((int) fa[(size_t) (j0 + r) * a_stride + i] - (int) fb[(size_t) (j0 + r) * b_stride + i]);
    }
  }
  // CTA reduction: warp shuffles, then one atomic per CTA
  __shared__ uint32_t part[VF_TX * VF_TY / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync (0xffffffffu, acc, o);
  const int tid = threadIdx.y * VF_TX + threadIdx.x;
  if ((tid & 31) == 0) part[tid >> 5] = acc;
  __syncthreads ();
  if (tid == 0) {
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < VF_TX * VF_TY / 32; k++) s += part[k];
    if (s) atomicAdd (sums + blockIdx.z, s);
  }
}

// ---- smooth (gst/smooth/gstsmooth.c:131-176, its own plugin; same boundary, I420 planes) -----------------
// Adaptive box filter: the output is the mean of the window samples within +-tolerance of the reference sample
// (plus the reference sample itself, counted once more), integer division. The reference's loop has two
// quirks this kernel reproduces (smooth_geometry): the row pointers are advanced at the END of an iteration to
// row y, so iteration y reads its reference sample from row y-1 and writes row y-1 with the window of row y -
// output row r is computed with rows [max(0, r-fs), r+fs+3) (clipped as the loop clips them), row 0 is written
// twice (the second time wins) and the LAST row is never written; the columns are the symmetric [x-fs, x+fs].
// Unlike the other filters here this one is bound by integer issue, not by HBM: (2fs+1)(2fs+3) = 63 compares
// per sample at the default filter-size 3.
constexpr int SM_TW = 64, SM_TH = 16, SM_FS_MAX = 8;

// window rows [ra, rb) of loop iteration y (fy1 / fy2 of the reference divided by the stride)
__host__ __device__ inline void smooth_rows (int y, int height, int fs, int *ra, int *rb) {
  const int a = y - (fs + 1);
  *ra = a > 0 ? a : 0;
  const int lim = height - (fs + 1) > 0 ? height - (fs + 1) : 0;
  *rb = (fs + 1 < height ? fs + 1 : height) + (y + 1 < lim ? y + 1 : lim);
}

__global__ void __launch_bounds__ (256)
smooth_kernel (const uint8_t *src, int src_stride, size_t src_fs, uint8_t *dst, int dst_stride, size_t dst_fs,
    int width, int height, int atol, int fs /* >= 0 */, int empty_window)
{
  __shared__ uint8_t tile[SM_TH + 2 * SM_FS_MAX + 2][SM_TW + 2 * SM_FS_MAX];
  const uint8_t *s = src + (size_t) blockIdx.z * src_fs;
  uint8_t *d = dst + (size_t) blockIdx.z * dst_fs;
  const int x0 = blockIdx.x * SM_TW, r0 = blockIdx.y * SM_TH;
  const int tid = threadIdx.y * 64 + threadIdx.x;
  const int trows = SM_TH + 2 * fs + 2, tcols = SM_TW + 2 * fs;
  for (int i = tid; i < trows * tcols; i += 256) {           // rows r0-fs .. r0+TH+fs+1, columns x0-fs .. x0+TW+fs-1
    const int tr = i / tcols, tc = i % tcols;
    const int gr = r0 - fs + tr, gc = x0 - fs + tc;
    tile[tr][tc] = (gr >= 0 && gr < height && gc >= 0 && gc < width) ? s[(size_t) gr * src_stride + gc] : (uint8_t) 0;
  }
  __syncthreads ();
  const int x = x0 + threadIdx.x;
  const int last = height >= 2 ? height - 2 : 0;             // last row the reference writes
  if (x >= width) return;
  const int c0 = max (x - fs, 0), c1 = empty_window ? c0 : min (x + fs + 1, width);
#pragma unroll 1
  for (int k = 0; k < SM_TH / 4; k++) {
    const int r = r0 + threadIdx.y * (SM_TH / 4) + k;
    if (r > last) break;
    int ra, rb;
    smooth_rows (height >= 2 ? r + 1 : 0, height, fs, &ra, &rb);
    const int ref = tile[r - r0 + fs][x - x0 + fs];
    int num = 1, sum = ref;
    for (int wr = ra; wr < rb; wr++) {
      const uint8_t *row = tile[wr - r0 + fs] + (fs - x0);
      for (int wc = c0; wc < c1; wc++) {
        const int akt = row[wc];
        if (abs (akt - ref) < atol) { num++; sum += akt; }
      }
    }
    d[(size_t) r * dst_stride + x] = (uint8_t) (sum / num);
  }
}

// The same filter with four window samples per instruction (round 2; the scalar kernel above spends its time in 63
// LDS.U8 + compare + add chains per sample: 0.01 of the HBM peak). A thread owns the 4 output pixels of one aligned
// 32-bit word. The tile is kept in shared memory in FOUR byte-shifted copies (copy s, word i = bytes 4i+s .. 4i+s+3 of
// the row), so that any 4-byte group of window samples, at any byte offset, is ONE aligned LDS.32 - no funnel shifts on
// the ALU pipe, which is the half-rate pipe that bounds this kernel (ncu: alu 83 %, issue 82 %). A group then costs
//   VABSDIFF4          |sample - ref| per byte (ref replicated in the four lanes)
//   LOP3, IADD, LOP3   per-byte unsigned "< tolerance" without cross-byte carries, flag in bit 7 of each byte; lanes
//                      outside the window / the frame carry a threshold of 0 and never pass
//   IDP.4A x 2         sum += samples . flags and count += 1 . flags (both scaled by 128: the flags are 0x80)
// i.e. ~2 instructions per sample instead of ~6. MODE 0: tolerance <= 128 (the default is 8): a < t per byte is bit 7 of
// ((t + 127) - (a & 127)) & ~a, no carry can cross a byte; MODE 1: any tolerance <= 255 (one more LOP3); MODE 2:
// |tolerance| >= 256 admits every sample of the window. The final sum / count (count <= 196, sum < 2^16) is
// floor ((sum + 0.5) * rcp (count)) in fp32: the +0.5 keeps the approximate reciprocal's error (<= 2 ulp) away from
// every integer boundary (the nearest one is 0.5 / count >= 1.5e-3 away) - checked against integer division for every
// (count, sum) and every reciprocal within +-3 ulp in tests/test_host_logic_cpu.py::test_smooth_division_trick_is_exact.
// The window geometry is the scalar kernel's (smooth_rows; the reference's quirks are listed there).
// floor (total / num) for num in [1, 324], total <= 255 * num (see the kernel comment): one MUFU.RCP (rcp.approx, within
// 2 ulp of 1 / num) instead of the ~20 instructions and branches of an exact reciprocal or an integer division
__device__ __forceinline__ uint32_t smooth_div (uint32_t total, uint32_t num) {
  float r;
  asm ("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"((float) num));
  return (uint32_t) (((float) total + 0.5f) * r);
}

template <int FS> struct SmoothGeom {
  static constexpr int HW = (FS + 3) / 4;                      // halo words each side of the thread's own word
  static constexpr int NW = (2 * FS + 1 + 3) / 4;              // 4-sample groups per pixel and window row
};
constexpr int SMD_TWW = 32, SMD_TH = 32;                       // tile: 32 words (128 pixels) x 32 rows, 256 threads x 4 rows

template <int FS, int MODE>
__global__ void __launch_bounds__ (256)
smooth_dp4a_kernel (const uint8_t *src, int src_stride, size_t src_fs, uint8_t *dst, int dst_stride, size_t dst_fs,
    int width, int height, int atol /* 0 .. 256 */)
{
  typedef SmoothGeom<FS> G;
  constexpr int TROWS = SMD_TH + 2 * FS + 2, TWORDS = SMD_TWW + 2 * G::HW + 1;
  extern __shared__ uint32_t smooth_tile[];                    // [4 shifts][TROWS][TWORDS]
  uint32_t (*tile)[TROWS][TWORDS] = reinterpret_cast<uint32_t (*)[TROWS][TWORDS]> (smooth_tile);
  const uint8_t *s = src + (size_t) blockIdx.z * src_fs;
  uint8_t *d = dst + (size_t) blockIdx.z * dst_fs;
  const int xw0 = blockIdx.x * SMD_TWW, r0 = blockIdx.y * SMD_TH;       // first word column / first row of the tile
  const int tid = threadIdx.y * 32 + threadIdx.x;
  const int nwords = (width + 3) >> 2;                                   // words that hold samples (the last one possibly partly)
  for (int i = tid; i < TROWS * TWORDS; i += 256) {                     // rows r0-FS .., word columns xw0-HW ..
    const int tr = i / TWORDS, tc = i % TWORDS;
    const int gr = r0 - FS + tr, gw = xw0 - G::HW + tc;
    uint32_t v = 0, n = 0;
    if (gr >= 0 && gr < height) {
      if (gw >= 0 && gw < nwords) v = ldg_u32 (s + (size_t) gr * src_stride + 4 * gw);
      if (gw + 1 >= 0 && gw + 1 < nwords) n = ldg_u32 (s + (size_t) gr * src_stride + 4 * (gw + 1));
    }
    tile[0][tr][tc] = v;
    tile[1][tr][tc] = __funnelshift_r (v, n, 8);
    tile[2][tr][tc] = __funnelshift_r (v, n, 16);
    tile[3][tr][tc] = __funnelshift_r (v, n, 24);
  }
  __syncthreads ();
  const int xw = xw0 + threadIdx.x, x0 = 4 * xw;
  if (x0 >= width) return;
  // per pixel p and group k: the threshold word - the tolerance in the lanes whose column lies in the pixel's window
  // [x - FS, x + FS] and in the frame, 0 elsewhere (a lane with threshold 0 never passes). MODE 0 stores t + 127 per
  // byte, MODE 2 the flag itself.
  uint32_t thr[4][G::NW];
  const uint32_t at = MODE == 2 ? 0x80u : MODE == 0 ? (uint32_t) atol + 127u : (uint32_t) atol;
#pragma unroll
  for (int p = 0; p < 4; p++)
#pragma unroll
    for (int k = 0; k < G::NW; k++) {
      uint32_t t = MODE == 0 ? 0x7f7f7f7fu : 0u;
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const int c = -FS + 4 * k + j, col = x0 + p + c;
        if (c <= FS && col >= 0 && col < width) t = (t & ~(0xffu << (8 * j))) | (at << (8 * j));
      }
      thr[p][k] = t;
    }
  const int last = height >= 2 ? height - 2 : 0;                         // last row the reference writes
#pragma unroll 1
  for (int q = 0; q < SMD_TH / 8; q++) {
    const int r = r0 + threadIdx.y * (SMD_TH / 8) + q;
    if (r > last) break;
    int ra, rb;
    smooth_rows (height >= 2 ? r + 1 : 0, height, FS, &ra, &rb);
    const uint32_t refw = tile[0][r - r0 + FS][threadIdx.x + G::HW];
    uint32_t ref4[4];
    uint32_t sum[4] = { 0, 0, 0, 0 }, cnt[4] = { 0, 0, 0, 0 };
#pragma unroll
    for (int p = 0; p < 4; p++) ref4[p] = ((refw >> (8 * p)) & 0xffu) * 0x01010101u;
#pragma unroll 1
    for (int wr = ra; wr < rb; wr++) {
      const int tr = wr - r0 + FS;
#pragma unroll
      for (int p = 0; p < 4; p++)
#pragma unroll
        for (int k = 0; k < G::NW; k++) {
          constexpr int B0 = 4 * G::HW - FS;                             // first byte of pixel 0's window, counted from the thread's word 0
          const int b = p + B0 + 4 * k;                                  // compile-time after unrolling
          const uint32_t v = tile[b & 3][tr][threadIdx.x + (b >> 2)];
          uint32_t f;
          if (MODE == 2) f = thr[p][k];
          else {
            const uint32_t a = __vabsdiffu4 (v, ref4[p]);
            if (MODE == 0) f = (thr[p][k] - (a & 0x7f7f7f7fu)) & ~a & 0x80808080u;
            else {
              const uint32_t t = thr[p][k], h = (~a & 0x7f7f7f7fu) + (t & 0x7f7f7f7fu);
              f = ((~a & t) | (~(a ^ t) & h)) & 0x80808080u;
            }
          }
          sum[p] = __dp4a (v, f, sum[p]);
          cnt[p] = __dp4a (0x01010101u, f, cnt[p]);
        }
    }
    uint32_t out = 0;
#pragma unroll
    for (int p = 0; p < 4; p++) {
      const uint32_t total = (sum[p] >> 7) + (ref4[p] & 0xffu), num = (cnt[p] >> 7) + 1u;
      out |= smooth_div (total, num) << (8 * p);
    }
    uint8_t *o = d + (size_t) r * dst_stride + x0;
    if (x0 + 4 <= width) *reinterpret_cast<uint32_t *> (o) = out;
    else for (int p = 0; x0 + p < width; p++) o[p] = (uint8_t) (out >> (8 * p));
  }
}

// ---- videoanalyse (gst/videosignal/gstvideoanalyse.c:206-236) --------------------------------------------
// The reference walks the luma plane twice: sum, then sum of (avg - d)^2 with avg = sum / (w*h) in int. The second
// sum is an exact integer identity of the first two moments, N*avg^2 - 2*avg*sum + sum(d^2), so one pass that
// accumulates sum(d) and sum(d^2) suffices (dp4a: four samples per instruction each); 64-bit atomics per CTA.
template <int WORDS>
__global__ void __launch_bounds__ (VF_TX * VF_TY)
luma_sums_kernel (const uint8_t *a, int a_stride, size_t a_fs, int width, int height, unsigned long long *sums)
{
  const int w0 = (blockIdx.x * VF_TX + threadIdx.x) * WORDS;
  const int j0 = (blockIdx.y * VF_TY + threadIdx.y) * VF_ROWS;
  const int nwords = width >> 2;
  const uint8_t *fa = a + (size_t) blockIdx.z * a_fs;
  uint32_t s1 = 0, s2 = 0;                                   // <= 64 samples per thread: 64 * 65025 fits
  if (w0 < nwords) {
    uint32_t va[VF_ROWS][WORDS];
#pragma unroll
    for (int r = 0; r < VF_ROWS; r++) {
#pragma unroll
      for (int k = 0; k < WORDS; k++) va[r][k] = 0;
      if (j0 + r >= height) continue;
      const uint32_t *pa = reinterpret_cast<const uint32_t *> (fa + (size_t) (j0 + r) * a_stride) + w0;
      if (WORDS == 4 && w0 + 4 <= nwords) { const uint4 q = ld_stream_v4 (pa); va[r][0] = q.x; va[r][1] = q.y; va[r][2] = q.z; va[r][3] = q.w; }
      else
#pragma unroll
        for (int k = 0; k < WORDS; k++) if (w0 + k < nwords) va[r][k] = ldg_u32 (pa + k);
    }
#pragma unroll
    for (int r = 0; r < VF_ROWS; r++)
#pragma unroll
      for (int k = 0; k < WORDS; k++) { s1 = __dp4a (va[r][k], 0x01010101u, s1); s2 = __dp4a (va[r][k], va[r][k], s2); }
  }
  if (blockIdx.x == 0 && threadIdx.x < (width & 3)) {
    const int i = (nwords << 2) + threadIdx.x;
    for (int r = 0; r < VF_ROWS; r++) {
      if (j0 + r >= height) break;
      const uint32_t v = fa[(size_t) (j0 + r) * a_stride + i];
      s1 += v; s2 += v * v;
    }
  }
  __shared__ unsigned long long part[2][VF_TX * VF_TY / 32];
  unsigned long long t1 = s1, t2 = s2;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { t1 += __shfl_xor_sync (0xffffffffu, t1, o); t2 += __shfl_xor_sync (0xffffffffu, t2, o); }
  const int tid = threadIdx.y * VF_TX + threadIdx.x;
  if ((tid & 31) == 0) { part[0][tid >> 5] = t1; part[1][tid >> 5] = t2; }
  __syncthreads ();
  if (tid < 2) {
    unsigned long long s = 0;
#pragma unroll
    for (int k = 0; k < VF_TX * VF_TY / 32; k++) s += part[tid][k];
    if (s) atomicAdd (sums + 2 * (size_t) blockIdx.z + tid, s);
  }
}

bool aligned16 (const void *p, size_t a, size_t b) { return ((uintptr_t) p) % 16 == 0 && a % 16 == 0 && b % 16 == 0; }
bool aligned4 (const void *p, size_t a, size_t b) { return ((uintptr_t) p) % 4 == 0 && a % 4 == 0 && b % 4 == 0; }

dim3 vf_grid (int width, int height, int nframes, int words) {
  const int nwords = width >> 2;
  int gx = (nwords + VF_TX * words - 1) / (VF_TX * words);
  if (gx < 1) gx = 1;                                        // width < 4: only the bytewise tail
  return dim3 (gx, (height + VF_TY * VF_ROWS - 1) / (VF_TY * VF_ROWS), nframes);
}

}  // namespace

B200VF_API int b200vf_zebrastripe_y_threshold (int threshold) {
  return 16 + (int) floor (0.5 + 2.19 * threshold);          // gstzebrastripe.c:150-151
}

B200VF_API int b200vf_zebrastripe (b200vf_ctx *ctx, uint8_t *d_luma, int pixel_stride, int row_stride, size_t frame_stride,
    int nframes, int width, int height, int y_threshold, int t, void *stream)
{
  B200VF_REQUIRE (ctx && d_luma && width > 0 && height > 0 && nframes > 0, B200VF_E_INVAL, "zebrastripe: bad argument");
  B200VF_REQUIRE (pixel_stride == 1 || pixel_stride == 2 || pixel_stride == 4, B200VF_E_UNSUPPORTED,
      "zebrastripe: pixel stride %d (1 planar / NV12, 2 YUY2 / UYVY, 4 AYUV)", pixel_stride);
  B200VF_REQUIRE (row_stride >= width * pixel_stride - (pixel_stride - 1), B200VF_E_INVAL, "zebrastripe: row stride");
  B200VF_REQUIRE (height <= 65535 * VF_TY * VF_ROWS && nframes <= 65535, B200VF_E_UNSUPPORTED, "zebrastripe: grid limits");
  cudaStream_t s = b200vf_stream (ctx, stream);
  const dim3 block (VF_TX, VF_TY);
  if (pixel_stride == 1 && aligned4 (d_luma, (size_t) row_stride, frame_stride)) {
    if (aligned16 (d_luma, (size_t) row_stride, frame_stride))
      zebra_planar_kernel<4><<<vf_grid (width, height, nframes, 4), block, 0, s>>> (d_luma, row_stride, frame_stride, width, height, y_threshold, t);
    else
      zebra_planar_kernel<1><<<vf_grid (width, height, nframes, 1), block, 0, s>>> (d_luma, row_stride, frame_stride, width, height, y_threshold, t);
    return b200vf_launched (ctx, "zebrastripe_planar");
  }
  const dim3 grid ((width + VF_TX - 1) / VF_TX, (height + VF_TY * VF_ROWS - 1) / (VF_TY * VF_ROWS), nframes);
  zebra_packed_kernel<<<grid, block, 0, s>>> (d_luma, pixel_stride, row_stride, frame_stride, width, height, y_threshold, t);
  return b200vf_launched (ctx, "zebrastripe_packed");
}

B200VF_API int b200vf_videodiff_luma (b200vf_ctx *ctx, const uint8_t *d_old, int old_stride, size_t old_frame_stride,
    const uint8_t *d_new, int new_stride, size_t new_frame_stride, uint8_t *d_out, int out_stride, size_t out_frame_stride,
    int width, int height, int nframes, int threshold, int t, void *stream)
{
  B200VF_REQUIRE (ctx && d_old && d_new && d_out && width > 0 && height > 0 && nframes > 0, B200VF_E_INVAL, "videodiff: bad argument");
  B200VF_REQUIRE (old_stride >= width && new_stride >= width && out_stride >= width, B200VF_E_INVAL, "videodiff: row stride");
  B200VF_REQUIRE (aligned4 (d_old, (size_t) old_stride, old_frame_stride) && aligned4 (d_new, (size_t) new_stride, new_frame_stride) &&
      aligned4 (d_out, (size_t) out_stride, out_frame_stride), B200VF_E_INVAL,
      "videodiff: planes and pitches must be 4-byte aligned (GstVideoInfo rounds luma pitches up to 4)");
  B200VF_REQUIRE (height <= 65535 * VF_TY * VF_ROWS && nframes <= 65535, B200VF_E_UNSUPPORTED, "videodiff: grid limits");
  cudaStream_t s = b200vf_stream (ctx, stream);
  const dim3 block (VF_TX, VF_TY);
  const bool v16 = aligned16 (d_old, (size_t) old_stride, old_frame_stride) && aligned16 (d_new, (size_t) new_stride, new_frame_stride) &&
      aligned16 (d_out, (size_t) out_stride, out_frame_stride);
  if (v16)
    videodiff_kernel<4><<<vf_grid (width, height, nframes, 4), block, 0, s>>> (d_old, old_stride, old_frame_stride, d_new, new_stride,
        new_frame_stride, d_out, out_stride, out_frame_stride, width, height, threshold, t);
  else
    videodiff_kernel<1><<<vf_grid (width, height, nframes, 1), block, 0, s>>> (d_old, old_stride, old_frame_stride, d_new, new_stride,
        new_frame_stride, d_out, out_stride, out_frame_stride, width, height, threshold, t);
  return b200vf_launched (ctx, "videodiff_luma");
}

B200VF_API int b200vf_sad_u8 (b200vf_ctx *ctx, const uint8_t *d_a, int a_stride, size_t a_frame_stride, const uint8_t *d_b,
    int b_stride, size_t b_frame_stride, int width, int height, int nframes, uint32_t *d_sums, void *stream)
{
  B200VF_REQUIRE (ctx && d_a && d_b && d_sums && width > 0 && height > 0 && nframes > 0, B200VF_E_INVAL, "sad_u8: bad argument");
  B200VF_REQUIRE (a_stride >= width && b_stride >= width, B200VF_E_INVAL, "sad_u8: row stride");
  B200VF_REQUIRE (aligned4 (d_a, (size_t) a_stride, a_frame_stride) && aligned4 (d_b, (size_t) b_stride, b_frame_stride), B200VF_E_INVAL,
      "sad_u8: planes and pitches must be 4-byte aligned");
  B200VF_REQUIRE (height <= 65535 * VF_TY * VF_ROWS && nframes <= 65535, B200VF_E_UNSUPPORTED, "sad_u8: grid limits");
  cudaStream_t s = b200vf_stream (ctx, stream);
  B200VF_CHECK_CUDA (cudaMemsetAsync (d_sums, 0, sizeof (uint32_t) * (size_t) nframes, s));
  const dim3 block (VF_TX, VF_TY);
  if (aligned16 (d_a, (size_t) a_stride, a_frame_stride) && aligned16 (d_b, (size_t) b_stride, b_frame_stride))
    sad_kernel<4><<<vf_grid (width, height, nframes, 4), block, 0, s>>> (d_a, a_stride, a_frame_stride, d_b, b_stride, b_frame_stride,
        width, height, d_sums);
  else
    sad_kernel<1><<<vf_grid (width, height, nframes, 1), block, 0, s>>> (d_a, a_stride, a_frame_stride, d_b, b_stride, b_frame_stride,
        width, height, d_sums);
  return b200vf_launched (ctx, "sad_u8");
}

// smooth_filter on one plane (luma, or a chroma plane when luma-only is off). Rows 0 .. height-2 of the
// destination are written (row 0 only, for a one-row plane); the last row keeps what the buffer held, as in the
// reference. (lower - akt) * (upper - akt) < 0 is |akt - ref| < |tolerance| for every tolerance that does not
// overflow the reference's int product; |tolerance| > 255 admits every sample.
B200VF_API int b200vf_smooth_plane (b200vf_ctx *ctx, const uint8_t *d_src, int src_stride, size_t src_frame_stride,
    uint8_t *d_dst, int dst_stride, size_t dst_frame_stride, int width, int height, int nframes, int tolerance,
    int filtersize, void *stream)
{
  B200VF_REQUIRE (ctx && d_src && d_dst && width > 0 && height > 0 && nframes > 0, B200VF_E_INVAL, "smooth: bad argument");
  B200VF_REQUIRE (src_stride >= width && dst_stride >= width, B200VF_E_INVAL, "smooth: row stride");
  B200VF_REQUIRE (filtersize <= SM_FS_MAX, B200VF_E_UNSUPPORTED, "smooth: filter-size %d (supported: <= %d)", filtersize, SM_FS_MAX);
  B200VF_REQUIRE ((height + SM_TH - 1) / SM_TH <= 65535 && nframes <= 65535, B200VF_E_UNSUPPORTED, "smooth: grid limits");
  cudaStream_t s = b200vf_stream (ctx, stream);
  long long at = tolerance < 0 ? -(long long) tolerance : (long long) tolerance;
  const int atol = at > 256 ? 256 : (int) at;
  // a negative filter-size leaves the window empty (fx1 >= fx2 in the reference's loop): the output is the reference sample
  const int fs = filtersize < 0 ? 0 : filtersize;
  // four samples per instruction when the rows are word-aligned and the window fits the packed counters
  const bool packed = filtersize >= 0 && filtersize <= 6 && src_stride % 4 == 0 && dst_stride % 4 == 0 && ((uintptr_t) d_src) % 4 == 0 &&
      ((uintptr_t) d_dst) % 4 == 0 && src_frame_stride % 4 == 0 && dst_frame_stride % 4 == 0 && !getenv ("B200VF_SMOOTH_SCALAR");
  if (packed) {
    const dim3 blk (32, 8), grd ((width + 4 * SMD_TWW - 1) / (4 * SMD_TWW), (height + SMD_TH - 1) / SMD_TH, nframes);
    B200VF_REQUIRE (grd.y <= 65535, B200VF_E_UNSUPPORTED, "smooth: grid limits");
    const int mode = atol > 255 ? 2 : atol <= 128 ? 0 : 1;
#define SMOOTH_LAUNCH(F, M) { \
      const int smem = 4 * (SMD_TH + 2 * F + 2) * (SMD_TWW + 2 * SmoothGeom<F>::HW + 1) * 4; \
      int rcs = b200vf_func_smem (ctx, (const void *) smooth_dp4a_kernel<F, M>, smem); \
      if (rcs) return rcs; \
      smooth_dp4a_kernel<F, M><<<grd, blk, smem, s>>> (d_src, src_stride, src_frame_stride, d_dst, dst_stride, dst_frame_stride, width, height, atol); }
#define SMOOTH_CASE(F) case F: if (mode == 0) SMOOTH_LAUNCH (F, 0) else if (mode == 1) SMOOTH_LAUNCH (F, 1) else SMOOTH_LAUNCH (F, 2) break;
    switch (fs) { SMOOTH_CASE (0) SMOOTH_CASE (1) SMOOTH_CASE (2) SMOOTH_CASE (3) SMOOTH_CASE (4) SMOOTH_CASE (5) default: SMOOTH_CASE (6) }
#undef SMOOTH_CASE
#undef SMOOTH_LAUNCH
    return b200vf_launched (ctx, "smooth_dp4a");
  }
  const dim3 block (64, 4), grid ((width + SM_TW - 1) / SM_TW, (height + SM_TH - 1) / SM_TH, nframes);
  smooth_kernel<<<grid, block, 0, s>>> (d_src, src_stride, src_frame_stride, d_dst, dst_stride, dst_frame_stride, width, height, atol, fs,
      filtersize < 0 ? 1 : 0);
  return b200vf_launched (ctx, "smooth");
}

// videoanalyse: the two moments of each luma plane; d_sums[2f] = sum, d_sums[2f+1] = sum of squares (device).
B200VF_API int b200vf_luma_moments (b200vf_ctx *ctx, const uint8_t *d_luma, int stride, size_t frame_stride, int width, int height,
    int nframes, uint64_t *d_sums, void *stream)
{
  B200VF_REQUIRE (ctx && d_luma && d_sums && width > 0 && height > 0 && nframes > 0, B200VF_E_INVAL, "luma_moments: bad argument");
  B200VF_REQUIRE (stride >= width && aligned4 (d_luma, (size_t) stride, frame_stride), B200VF_E_INVAL,
      "luma_moments: plane and pitch must be 4-byte aligned, pitch >= width");
  B200VF_REQUIRE (height <= 65535 * VF_TY * VF_ROWS && nframes <= 65535, B200VF_E_UNSUPPORTED, "luma_moments: grid limits");
  cudaStream_t s = b200vf_stream (ctx, stream);
  B200VF_CHECK_CUDA (cudaMemsetAsync (d_sums, 0, 2 * sizeof (uint64_t) * (size_t) nframes, s));
  const dim3 block (VF_TX, VF_TY);
  unsigned long long *out = reinterpret_cast<unsigned long long *> (d_sums);
  if (aligned16 (d_luma, (size_t) stride, frame_stride))
    luma_sums_kernel<4><<<vf_grid (width, height, nframes, 4), block, 0, s>>> (d_luma, stride, frame_stride, width, height, out);
  else
    luma_sums_kernel<1><<<vf_grid (width, height, nframes, 1), block, 0, s>>> (d_luma, stride, frame_stride, width, height, out);
  return b200vf_launched (ctx, "luma_moments");
}

// gst_video_analyse_planar (:219-220, :232): the element's two numbers from the moments, in the reference's
// arithmetic: avg = sum / (w*h) in integers; luma_average = sum / (255.0 * w * h); luma_variance =
// sum ((avg - d)^2) / (255.0 * 255.0 * w * h), the integer sum via N*avg^2 - 2*avg*sum + sum(d^2) (exact).
B200VF_API int b200vf_videoanalyse_finish (uint64_t sum, uint64_t sum_sq, int width, int height, double *luma_average,
    double *luma_variance)
{
  B200VF_REQUIRE (width > 0 && height > 0 && luma_average && luma_variance, B200VF_E_INVAL, "videoanalyse_finish: bad argument");
  const uint64_t n = (uint64_t) (width * height);            // the reference's int product
  const int avg = (int) (sum / n);
  *luma_average = sum / (255.0 * width * height);
  const uint64_t var = n * (uint64_t) avg * (uint64_t) avg + sum_sq - 2ull * (uint64_t) avg * sum;
  *luma_variance = var / (255.0 * 255.0 * width * height);
  return B200VF_OK;
}

// gst_scene_change_transform_frame_ip, gstscenechange.c:196-236: the decision on the frame score
// (score = SAD / (width * height), :154). Host arithmetic on five doubles, statement for statement.
B200VF_API int b200vf_scenechange_reset (b200vf_scenechange_state *st) {
  B200VF_REQUIRE (st, B200VF_E_INVAL, "scenechange_reset: NULL");
  st->n_diffs = 0;
  for (int i = 0; i < B200VF_SC_N_DIFFS; i++) st->diffs[i] = 0.0;
  return B200VF_OK;
}
B200VF_API int b200vf_scenechange_update (b200vf_scenechange_state *st, double score, int *change_out) {
  B200VF_REQUIRE (st && change_out, B200VF_E_INVAL, "scenechange_update: NULL");
  const int N = B200VF_SC_N_DIFFS;
  for (int i = 0; i < N - 1; i++) st->diffs[i] = st->diffs[i + 1];     // memmove (:200-201)
  st->diffs[N - 1] = score;
  st->n_diffs++;
  double score_min = st->diffs[0], score_max = st->diffs[0];
  for (int i = 1; i < N - 1; i++) {
    score_min = score_min < st->diffs[i] ? score_min : st->diffs[i];     // MIN (score_min, diffs[i])
    score_max = score_max > st->diffs[i] ? score_max : st->diffs[i];     // MAX (score_max, diffs[i])
  }
  const double threshold = 1.8 * score_max - 0.8 * score_min;
  int change;
  if (st->n_diffs > (N - 1)) {
    if (score < 5) change = 0;
    else if (score / threshold < 1.0) change = 0;
    else if ((score > 30) && (score / st->diffs[N - 2] > 1.4)) change = 1;
    else if (score / threshold > 2.3) change = 1;
    else if (score > 50) change = 1;
    else change = 0;
  } else change = 0;
  if (change) b200vf_scenechange_reset (st);
  *change_out = change;
  return B200VF_OK;
}
