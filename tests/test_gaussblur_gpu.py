"""gaussianblur parity through the C-ABI. exact=1 is bit-exact (separate fp32 mul/add in the
reference's tap order); exact=0 (FMA) must stay within 1 LSB of the u8 output."""
import numpy as np
import pytest

import frames

pytestmark = pytest.mark.gpu


def run(ctx, vf, frame, w, h, sigma, p0, exact=True):
    k, ks = vf.gauss_kernel(sigma)
    stride = frame.shape[1]
    d_src = ctx.upload(frame)
    d_dst = ctx.alloc(frame.size)
    ctx.gaussblur(d_src, d_dst, w, h, stride, p0, k, ks, exact=exact)
    return ctx.download(d_dst, frame.size).reshape(frame.shape)


@pytest.mark.parametrize("sigma", [-5, -1.2, 0, 0.3, 1.2, 5, 20])
@pytest.mark.parametrize("p0", [0, 1, 2, 3])
def test_sigmas_small_frames(ctx, vf, orc, rng, sigma, p0):
    for (w, h) in [(8, 8), (40, 30), (70, 66)]:      # 8x8 at sigma=5: both edges truncate at once; 70x66: > one tile
        fr = frames.random_u8(rng, h, 4 * w)
        got = run(ctx, vf, fr, w, h, sigma, p0)
        want = orc.gaussblur(fr, w, h, sigma, p0)
        assert np.array_equal(got, want), (sigma, p0, w, h, ctx.last_kernel(), np.abs(got.astype(int) - want).max(), np.argwhere(got != want)[:4])


@pytest.mark.parametrize("gth,ctas", [(8, 1), (8, 3), (16, 2), (32, 5), (64, 4)])
@pytest.mark.parametrize("sigma,p0", [(5, 1), (1.2, 0), (12.5, 3), (-5, 2)])
@pytest.mark.parametrize("w", [75, 76])          # rows of 300 B (pre-pass route) and of 304 B (16-byte aligned: TMA reads the frame itself)
def test_strip_walk_carries_rows_between_steps(ctx, vf, orc, rng, monkeypatch, gth, ctas, sigma, p0, w):
    """A CTA owns a contiguous range of (strip, step) units and MOVES the last 2*center fp32 rows of a step to
    the top of its tile for the next one. Few CTAs and short steps (tuning knobs of the library) make every
    case of that walk happen on a small frame: ranges that start in mid-strip, ranges that span strips and
    frames, halos longer than a step (moved in several batches), a last step shorter than the others."""
    monkeypatch.setenv("B200VF_GAUSS_GTH", str(gth))
    monkeypatch.setenv("B200VF_GAUSS_CTAS", str(ctas))
    h, n = 150, 2
    fr = frames.random_u8(rng, n * h, 4 * w)
    k, ks = vf.gauss_kernel(sigma)
    d_src = ctx.upload(fr)
    d_dst = ctx.alloc(fr.size + 64)
    ctx.gaussblur(d_src, d_dst, w, h, 4 * w, p0, k, ks, nframes=n)
    got = ctx.download(d_dst, fr.size).reshape(n, h, 4 * w)
    for i in range(n):                 # frames of a batch are independent: bytes past a frame read as 0 (D5 slack)
        want = orc.gaussblur(fr[i * h:(i + 1) * h], w, h, sigma, p0)
        assert np.array_equal(got[i], want), (i, np.argwhere(got[i] != want)[:6])
    # the knobs must not change a byte: same call with the default schedule
    monkeypatch.delenv("B200VF_GAUSS_GTH")
    monkeypatch.delenv("B200VF_GAUSS_CTAS")
    d_ref = ctx.alloc(fr.size + 64)
    ctx.gaussblur(d_src, d_ref, w, h, 4 * w, p0, k, ks, nframes=n)
    assert np.array_equal(ctx.download(d_ref, fr.size).reshape(n, h, 4 * w), got)


def test_mid_size_frame_many_units_per_cta(ctx, vf, orc, rng):
    """1024x1300: 33 strips x 21 steps = 693 units on <= 296 CTAs, the default schedule with carried rows"""
    w, h = 1024, 1300
    fr = frames.random_u8(rng, h, 4 * w)
    for sigma, p0 in [(5, 1), (2.0, 0)]:
        got = run(ctx, vf, fr, w, h, sigma, p0)
        want = orc.gaussblur(fr, w, h, sigma, p0)
        assert np.array_equal(got, want), (sigma, p0, np.argwhere(got != want)[:6])


@pytest.mark.parametrize("p0", [1, 2, 3])
@pytest.mark.parametrize("pad", [0, 16])
def test_aligned_view_equals_prepass_route_and_oracle(ctx, vf, orc, rng, monkeypatch, p0, pad):
    """16-byte aligned rows are blurred in place as aligned words (byte offset p0 handled at the frame's left and
    right edge only; unpadded rows fetch column w from the next row, padded rows find it in the padding);
    other pitches go through the realigning pre-pass. Both must give the reference's bytes."""
    w, h = 160, 140
    fr = frames.random_u8(rng, h, 4 * w + pad)
    for sigma in (5, 1.2):
        want = orc.gaussblur(fr, w, h, sigma, p0)
        got = run(ctx, vf, fr, w, h, sigma, p0)
        assert np.array_equal(got, want), ("aligned view", sigma, np.argwhere(got != want)[:6])
        monkeypatch.setenv("B200VF_GAUSS_PREPASS", "1")
        got2 = run(ctx, vf, fr, w, h, sigma, p0)
        monkeypatch.delenv("B200VF_GAUSS_PREPASS")
        assert np.array_equal(got2, want), ("pre-pass", sigma, np.argwhere(got2 != want)[:6])


@pytest.mark.parametrize("p0", [0, 1])
def test_widest_window_on_the_tiled_path(ctx, vf, orc, rng, p0):
    """|sigma| = 20: 101 taps, halo of 100 rows > the 64-row step (carried rows move in two batches), 158 KB of
    shared memory (one CTA per SM); sigma = -20 takes the IEEE-division / clamping epilogue (negative taps)"""
    w, h = 160, 140
    fr = frames.random_u8(rng, h, 4 * w)
    for sigma in (20, -20):
        got = run(ctx, vf, fr, w, h, sigma, p0)
        want = orc.gaussblur(fr, w, h, sigma, p0)
        assert np.array_equal(got, want), (sigma, p0, ctx.last_kernel(), np.argwhere(got != want)[:6])


def test_final_rounding_all_fp32(ctx):
    """(guint8) CLAMP (q + 0.5 [fp64], 0, 255) computed without fp64 (finish_bits / finish_u8): every fp32 q"""
    assert ctx.gauss_selftest_finish(0, 0xffffffff) == 0


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_many_random_pixels_are_bit_exact(ctx, vf, orc, seed):
    """Separate multiply/add vs a fused multiply-add differ in well under 1 % of the output bytes, so
    small frames can hide a contraction; 3 x 31k pixels x 3 sigmas make it visible."""
    rng = np.random.default_rng(seed)
    w, h = 192, 160
    fr = frames.random_u8(rng, h, 4 * w)
    for sigma, p0 in [(1.2, 1), (2.0, 0), (5.0, 1)]:
        got = run(ctx, vf, fr, w, h, sigma, p0)
        want = orc.gaussblur(fr, w, h, sigma, p0)
        assert np.array_equal(got, want), (sigma, p0, int((got != want).sum()), np.argwhere(got != want)[:4])


def test_kernel_taps_match_reference(vf, orc):
    for sigma in [-20, -5, -1.2, 0, 0.3, 1.2, 2.0, 5, 12.5, 20]:
        k, ks = vf.gauss_kernel(sigma)
        rk, rks = orc.gauss_kernel(sigma)
        assert np.array_equal(k.view(np.uint32), rk.view(np.uint32)), sigma
        assert np.array_equal(ks.view(np.uint32), rks.view(np.uint32)), sigma


def test_padded_stride(ctx, vf, orc, rng):
    w, h = 50, 20
    fr = frames.random_u8(rng, h, 4 * w + 8)
    for p0 in (0, 1):
        got = run(ctx, vf, fr, w, h, 1.2, p0)
        assert np.array_equal(got, orc.gaussblur(fr, w, h, 1.2, p0)), p0


def test_ayuv_default_sigma_640x480(ctx, vf, orc):
    w, h = 640, 480
    fr = frames.bars_rgbx(w, h)
    got = run(ctx, vf, fr, w, h, 1.2, 1)
    assert np.array_equal(got, orc.gaussblur(fr, w, h, 1.2, 1))


def test_fma_mode_within_one_lsb(ctx, vf, orc, rng):
    w, h = 128, 96
    fr = frames.random_u8(rng, h, 4 * w)
    got = run(ctx, vf, fr, w, h, 5, 1, exact=False)
    want = orc.gaussblur(fr, w, h, 5, 1)
    assert np.abs(got.astype(int) - want.astype(int)).max() <= 1      # tolerance: 1 LSB of the u8 output


def test_row_sharded_equals_whole(ctx, vf, orc, rng):
    """two row shards with 13(+1)-row halos reproduce the whole-frame result (global truncation rules)"""
    w, h, sigma, p0 = 64, 80, 5, 1
    fr = frames.random_u8(rng, h, 4 * w)
    want = orc.gaussblur(fr, w, h, sigma, p0)
    k, ks = vf.gauss_kernel(sigma)
    d_src = ctx.upload(fr)
    d_dst = ctx.alloc(fr.size)
    stride = 4 * w
    for (row0, rows) in [(0, 40), (40, 40)]:
        ctx.gaussblur(d_src.ptr + row0 * stride, d_dst.ptr + row0 * stride, w, rows, stride, p0, k, ks,
                      row0=row0, rows=rows, full_height=h)
    got = ctx.download(d_dst, fr.size).reshape(fr.shape)
    assert np.array_equal(got, want), np.argwhere(got != want)[:6]


@pytest.mark.parametrize("p0", [0, 1, 2, 3])
def test_row_shards_many_cuts_equal_whole(ctx, vf, p0):
    """every shard boundary reproduces the whole-frame bytes, including the channels of a row's last pixel that
    live in the next row (p0 > 0, unpadded rows): 4 seeds x 5 cuts, sigma 5 and 1.2; shards see center+1 halo rows"""
    w, h = 64, 120
    d_dst = ctx.alloc(h * 4 * w)
    d_ref = ctx.alloc(h * 4 * w)
    for seed in range(4):
        fr = frames.random_u8(np.random.default_rng(100 + seed), h, 4 * w)
        d_src = ctx.upload(fr)
        for sigma in (5, 1.2):
            k, ks = vf.gauss_kernel(sigma)
            ctx.gaussblur(d_src, d_ref, w, h, 4 * w, p0, k, ks)
            want = ctx.download(d_ref, fr.size)
            for cut in (16, 40, 61, 64, 100):
                for (row0, rows) in [(0, cut), (cut, h - cut)]:
                    ctx.gaussblur(d_src.ptr + row0 * 4 * w, d_dst.ptr + row0 * 4 * w, w, rows, 4 * w, p0, k, ks,
                                  row0=row0, rows=rows, full_height=h)
                got = ctx.download(d_dst, fr.size)
                assert np.array_equal(got, want), (seed, sigma, cut, np.argwhere(got.reshape(h, -1) != want.reshape(h, -1))[:6])


def test_4k_sigma5_band_exact(ctx, vf, orc, rng):
    """BASELINE.json configs[2] geometry: 3840x2160, sigma=5. The CPU oracle needs ~1.7 s/frame at 4K,
    so compare a 256-row band that contains the top edge exactly, plus a size-independent property
    on the whole frame: a constant image stays constant."""
    w, h = 3840, 2160
    fr = frames.random_u8(rng, h, 4 * w)
    got = run(ctx, vf, fr, w, h, 5, 1)
    band = 256
    want = orc.gaussblur(fr[: band + 13], w, band + 13, 5, 1)        # rows < band are unaffected by the cut at band+13
    assert np.array_equal(got[:band], want[:band])
    const = np.full((h, 4 * w), 93, np.uint8)
    out = run(ctx, vf, const, w, h, 5, 0)                # p0 = 0: no channel reaches into the zero slack (D5)
    assert (out == 93).all()


def test_reciprocal_division_is_ieee_exact(ctx, vf):
    """The blur divides by per-column/row constants with RN(1/b) and two FMA corrections when every tap is
    >= 0 and of ordinary size (gaussblur.cu div_rn). Under that condition a dividend is 0 or in
    [2^-64, 2^13): check EVERY fp32 of that range, for every divisor a frame edge can produce."""
    lo, hi = (127 - 64) << 23, (127 + 13) << 23
    assert ctx.gauss_selftest_div(1.0, 0, 1) == 0                      # a == +0
    ndiv = 0
    for sigma in [0.5, 1.2, 2.0, 5.0, 12.5, 20.0]:
        k, ks = vf.gauss_kernel(sigma)
        ws, c = len(k), len(k) // 2
        divisors = set()
        for i in range(c + 1):
            divisors.add(np.float32(np.float64(ks[ws - 1]) - (np.float64(ks[i - 1]) if i else 0.0)))
            divisors.add(np.float32(ks[c + i]))
        for b in sorted(divisors):
            assert ctx.gauss_selftest_div(b, lo, hi) == 0, (sigma, float(b))
            ndiv += 1
    assert ndiv > 100
    # the self-test does detect a differing quotient: 2^60 / 2^-100 overflows, where the reciprocal route gives NaN
    assert ctx.gauss_selftest_div(np.float32(2.0 ** -100), 187 << 23, (187 << 23) + 16) == 16


def test_one_fma_division_whitelist(ctx, vf):
    """gaussblur_stream.cuh divides interior outputs by the full kernel sum b with ONE fma: a / b == RN(a + a*e).
    Every whitelisted (b, e) pair is checked against IEEE division over EVERY dividend the blur can produce
    (a == 0 and 2^-64 <= a < 2^13, as in test_reciprocal_division_is_ieee_exact); the sums of the usual sigmas
    must be on the list (else the blur silently takes the slower general kernel)."""
    lo, hi = (127 - 64) << 23, (127 + 13) << 23
    listed = 0
    for bits in range(0x3f800000 - 8, 0x3f800000 + 9):
        b = np.array([bits], np.uint32).view(np.float32)[0]
        e = vf.gauss_div1_constant(b)
        if e is None:
            continue
        listed += 1
        assert ctx.gauss_selftest_div1(b, e, 0, 1) == 0, hex(bits)
        assert ctx.gauss_selftest_div1(b, e, lo, hi) == 0, (hex(bits), float(e))
    assert listed >= 5
    for sigma in [0.3, 0.5, 1.0, 1.2, 2.0, 3.3, 5.0]:
        k, ks = vf.gauss_kernel(sigma)
        assert vf.gauss_div1_constant(ks[-1]) is not None, (sigma, float(ks[-1]))
    # the self-test does see a wrong constant: e = 2^-24 fails for b = 1 - 2^-24 on power-of-two dividends (a tie)
    b = np.array([0x3f7fffff], np.uint32).view(np.float32)[0]
    assert ctx.gauss_selftest_div1(b, np.float32(2.0 ** -24), lo, hi) > 0


@pytest.mark.parametrize("sigma,p0,w,h", [(5, 1, 300, 200), (1.2, 1, 300, 200), (2.0, 2, 260, 97), (3.3, 0, 132, 70), (5, 3, 128, 64)])
def test_stream_kernel_is_taken_and_padded_windows_agree(ctx, vf, orc, rng, monkeypatch, sigma, p0, w, h):
    """exact blur with symmetric non-negative taps and 16-byte aligned rows runs the streaming kernel; a window padded
    with zero taps to the next instantiated size (and to a larger one) gives the same bytes as the oracle"""
    fr = frames.random_u8(rng, h, 4 * w)
    want = orc.gaussblur(fr, w, h, sigma, p0)
    monkeypatch.setenv("B200VF_GAUSS_STREAM_C", "4")     # (also lifts the size threshold below which the general kernel is kept)
    got = run(ctx, vf, fr, w, h, sigma, p0)
    assert ctx.last_kernel() in ("gaussblur_exact_stream", "gaussblur_lastcol", "gaussblur_gap_copy")
    assert np.array_equal(got, want), np.argwhere(got != want)[:6]
    monkeypatch.setenv("B200VF_GAUSS_STREAM_C", "13")
    got = run(ctx, vf, fr, w, h, sigma, p0)
    assert np.array_equal(got, want), ("padded to 13", np.argwhere(got != want)[:6])
    monkeypatch.delenv("B200VF_GAUSS_STREAM_C")
    monkeypatch.setenv("B200VF_GAUSS_NO_STREAM", "1")
    got = run(ctx, vf, fr, w, h, sigma, p0)
    assert ctx.last_kernel() in ("gaussblur_exact", "gaussblur_tail", "gaussblur_gap_copy")
    assert np.array_equal(got, want), ("general kernel", np.argwhere(got != want)[:6])


@pytest.mark.parametrize("sigma,c", [(0.5, 2), (1.0, 3), (1.2, 4), (2.0, 5), (2.3, 6), (2.7, 7), (3.0, 8), (3.5, 9), (3.9, 10), (4.3, 11),
                                     (4.7, 12), (5.0, 13)])
def test_every_instantiated_half_window_of_the_stream_kernel(ctx, vf, orc, rng, monkeypatch, sigma, c):
    """the streaming kernel exists for every half-window from 3 to 13; a window runs in the first one that holds it
    (zero taps in front) - and must give the oracle's bytes in every larger one too"""
    w, h = 260, 70
    fr = frames.random_u8(rng, h, 4 * w)
    k, ks = vf.gauss_kernel(sigma)
    assert len(k) == 2 * c + 1                          # center = ceil (2.5 * (gfloat) sigma), gstgaussblur.c:369
    want = orc.gaussblur(fr, w, h, sigma, 1)
    for C in range(3, 14):
        if C < c:
            continue
        monkeypatch.setenv("B200VF_GAUSS_STREAM_C", str(C))
        got = run(ctx, vf, fr, w, h, sigma, 1)
        assert ctx.last_kernel() in ("gaussblur_exact_stream", "gaussblur_lastcol", "gaussblur_gap_copy")
        assert np.array_equal(got, want), (sigma, C, np.argwhere(got != want)[:6])


@pytest.mark.parametrize("p0", [0, 1, 2, 3])
@pytest.mark.parametrize("w,h,pad", [(256, 100, 0), (257, 70, 12), (384, 64, 0), (131, 40, 4), (640, 33, 0)])
def test_stream_kernel_edges_and_shards(ctx, vf, orc, rng, monkeypatch, p0, w, h, pad):
    """the streaming kernel on small frames (size threshold lifted): widths that are multiples of the 128-column strip
    (aligned column w would be a strip of its own: the edge pixel columns come from gauss_lastcol_kernel), ragged widths,
    padded rows, few CTAs (ranges spanning strips and frames), and row shards against the whole-frame oracle"""
    if (4 * w + pad) % 16:
        pytest.skip("rows not 16-byte aligned: pre-pass route, general kernel")
    monkeypatch.setenv("B200VF_GAUSS_STREAM_C", "13")
    n = 2
    fr = frames.random_u8(rng, n * h, 4 * w + pad)
    stride = 4 * w + pad
    for sigma in (5, 1.2):
        k, ks = vf.gauss_kernel(sigma)
        d_src = ctx.upload(fr)
        d_dst = ctx.alloc(fr.size + 64)
        for ctas in (None, 3):
            if ctas:
                monkeypatch.setenv("B200VF_GAUSS_CTAS", str(ctas))
            ctx.gaussblur(d_src, d_dst, w, h, stride, p0, k, ks, nframes=n)
            got = ctx.download(d_dst, fr.size).reshape(n, h, stride)
            for i in range(n):
                want = orc.gaussblur(fr[i * h:(i + 1) * h], w, h, sigma, p0)
                assert np.array_equal(got[i][:, :4 * w + min(pad, p0)], want[:, :4 * w + min(pad, p0)]), (sigma, ctas, i, np.argwhere(got[i] != want)[:6])
        monkeypatch.delenv("B200VF_GAUSS_CTAS")
        # two row shards of frame 0 (center+1 halo rows are there: the shard reads them from the frame itself)
        one = fr[:h]
        want = orc.gaussblur(one, w, h, sigma, p0)
        d_one = ctx.upload(one)
        d_out = ctx.alloc(one.size + 64)
        cut = (h // 2) & ~1
        for (row0, rows) in [(0, cut), (cut, h - cut)]:
            ctx.gaussblur(d_one.ptr + row0 * stride, d_out.ptr + row0 * stride, w, rows, stride, p0, k, ks, row0=row0, rows=rows, full_height=h)
        got = ctx.download(d_out, one.size).reshape(h, stride)
        assert np.array_equal(got[:, :4 * w + min(pad, p0)], want[:, :4 * w + min(pad, p0)]), ("shards", sigma, np.argwhere(got != want)[:6])


def test_two_threads_blur_on_one_context(ctx, vf, orc, rng, monkeypatch):
    """ops are callable from any thread: two streaming threads blur byte-shifted frames on the same context, each on its
    own stream - the side-stream fork / join around the streaming kernel (edge columns, gap bytes) uses per-context
    events under a lock and must not cross the two calls' dependencies"""
    import threading
    import torch
    monkeypatch.setenv("B200VF_GAUSS_STREAM_C", "13")
    w, h, n = 384, 96, 6
    k, ks = vf.gauss_kernel(5.0)
    jobs = []
    for t in range(2):
        fr = [frames.random_u8(rng, h, 4 * w) for _ in range(n)]
        jobs.append({"frames": fr, "want": [orc.gaussblur(f, w, h, 5.0, 1) for f in fr], "got": [], "err": None,
                     "stream": torch.cuda.Stream()})

    def work(job):
        try:
            st = job["stream"].cuda_stream
            src = [ctx.upload(f) for f in job["frames"]]
            dst = [ctx.alloc(f.size + 64) for f in job["frames"]]
            for rep in range(3):
                for s, d in zip(src, dst):
                    ctx.gaussblur(s, d, w, h, 4 * w, 1, k, ks, stream=st)
            job["stream"].synchronize()
            job["got"] = [ctx.download(d, f.size).reshape(f.shape) for d, f in zip(dst, job["frames"])]
        except Exception as e:                                  # surfaced in the main thread
            job["err"] = e

    threads = [threading.Thread(target=work, args=(j,)) for j in jobs]
    [t.start() for t in threads]
    [t.join() for t in threads]
    for j in jobs:
        assert j["err"] is None, j["err"]
        for got, want in zip(j["got"], j["want"]):
            assert np.array_equal(got, want)
