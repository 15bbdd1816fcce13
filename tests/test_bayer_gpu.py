"""bayer2rgb parity: CUDA path (through the C-ABI) vs the oracle, bit-exact.
Matrix from SURVEY.md Appendix A: sizes x 4 patterns x 4 channel layouts x edge-exposing frames."""
import numpy as np
import pytest

import frames
from oracle import BAYER_FORMATS, RGB_OFFSETS

pytestmark = pytest.mark.gpu

SIZES = [(4, 3), (4, 4), (6, 5), (8, 8), (64, 33), (130, 20), (256, 17), (640, 480)]
LAYOUTS = ["RGBA", "BGRA", "ARGB", "ABGR"]


def run_gpu(ctx, src, w, h, fmt, out, nframes=1):
    stride = src.shape[-1]
    d_src = ctx.upload(src)
    d_dst = ctx.alloc(nframes * h * w * 4)
    ctx.bayer2rgb(d_src, stride, d_dst, 4 * w, w, h, BAYER_FORMATS[fmt], RGB_OFFSETS[out], nframes=nframes)
    res = ctx.download(d_dst).reshape(nframes, h, 4 * w)
    return res


@pytest.mark.parametrize("variant", ["direct", "auto"])
@pytest.mark.parametrize("w,h", SIZES)
def test_random_all_patterns_layouts(ctx, orc, rng, w, h, variant):
    ctx.set_variant(variant)
    try:
        src = frames.random_u8(rng, h, frames.round_up_4(w))
        for fmt in BAYER_FORMATS:
            for out in LAYOUTS:
                want = orc.bayer2rgb(src, w, h, fmt, out)
                got = run_gpu(ctx, src, w, h, fmt, out)[0]
                assert np.array_equal(got, want), (w, h, fmt, out, ctx.last_kernel(), np.argwhere(got != want)[:4])
    finally:
        ctx.set_variant("auto")


@pytest.mark.parametrize("w,h", [(8, 8), (130, 20), (512, 64)])
def test_edge_exposing_frames(ctx, orc, w, h):
    stride = frames.round_up_4(w)
    cases = [np.full((h, stride), 77, np.uint8), frames.checker_u8(h, stride)]
    for (y, x) in [(0, 0), (1, 1), (0, w - 1), (1, w - 2), (h - 1, 0), (h - 1, w - 1), (h - 2, w - 3), (h - 4, 3), (h // 2, w // 2)]:
        cases.append(frames.hot_pixel_u8(h, stride, y, x))
    for src in cases:
        for fmt in BAYER_FORMATS:
            want = orc.bayer2rgb(src, w, h, fmt, "RGBA")
            got = run_gpu(ctx, src, w, h, fmt, "RGBA")[0]
            assert np.array_equal(got, want), (w, h, fmt)


def test_batch_of_frames(ctx, orc, rng):
    w, h, n = 256, 48, 5
    src = rng.integers(0, 256, (n, h, w), dtype=np.uint8)
    got = run_gpu(ctx, src, w, h, "bggr", "BGRA", nframes=n)
    for i in range(n):
        assert np.array_equal(got[i], orc.bayer2rgb(src[i], w, h, "bggr", "BGRA")), i


def test_videotestsrc_like_640x480_bggr_bgrx(ctx, orc):
    """BASELINE.json configs[0]: bayer2rgb bggr -> BGRx 640x480."""
    w, h = 640, 480
    src = frames.mosaic_from_rgbx(frames.bars_rgbx(w, h), w, h, "bggr")
    got = run_gpu(ctx, src, w, h, "bggr", "BGRx")[0]
    assert np.array_equal(got, orc.bayer2rgb(src, w, h, "bggr", "BGRx"))


def test_domain_errors(ctx, vf):
    d = ctx.alloc(4096)
    for (w, h) in [(5, 8), (2, 8), (8, 2)]:
        with pytest.raises(vf.B200vfError) as e:
            ctx.bayer2rgb(d, 8, d, 4 * w, w, h, 0, (0, 1, 2))
        assert e.value.status == vf.E_INVAL
    with pytest.raises(vf.B200vfError) as e:
        ctx.bayer2rgb(d, 8, d, 32, 8, 8, 0, (0, 2, 1))      # not one of the four dispatched layouts
    assert e.value.status == vf.E_UNSUPPORTED


def test_4k_full_size_properties(ctx, orc, rng):
    """BASELINE.json configs[1] size: bit-exact on a band sample + size-independent properties."""
    w, h = 3840, 2160
    src = frames.random_u8(rng, h, w)
    got = run_gpu(ctx, src, w, h, "bggr", "RGBA")[0].reshape(h, w, 4)
    # alpha is 255 everywhere; the real samples pass through untouched (B at even/even, R at odd/odd, G elsewhere)
    assert (got[:, :, 3] == 255).all()
    assert np.array_equal(got[0::2, 0::2, 2], src[0::2, 0::2])
    assert np.array_equal(got[1::2, 1::2, 0], src[1::2, 1::2])
    assert np.array_equal(got[0::2, 1::2, 1], src[0::2, 1::2])
    assert np.array_equal(got[1::2, 0::2, 1], src[1::2, 0::2])
    # full-frame bit-exact vs the oracle (the C oracle takes ~20 ms at 4K)
    want = orc.bayer2rgb(src, w, h, "bggr", "RGBA").reshape(h, w, 4)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("variant", ["direct", "auto"])
@pytest.mark.parametrize("w,h,cuts", [(256, 96, [0, 32, 96]), (512, 200, [0, 66, 132, 200]), (64, 50, [0, 24, 50]),
                                      (3840, 2160, [0, 270, 540, 2160])])
def test_row_shards_equal_whole_frame(ctx, orc, rng, w, h, cuts, variant):
    """b200vf_bayer2rgb_shard: each shard reads its halo rows from the neighbouring rows of the same buffer;
    global edge rules (top mirror, bottom row h-4) apply with global row indices."""
    ctx.set_variant(variant)
    try:
        src = frames.random_u8(rng, h, w)
        want = orc.bayer2rgb(src, w, h, "grbg", "BGRA")
        d_src = ctx.upload(src)
        d_dst = ctx.alloc(h * w * 4)
        kernels = set()
        for r0, r1 in zip(cuts[:-1], cuts[1:]):
            ctx.bayer2rgb_shard(d_src.ptr + r0 * w, w, d_dst.ptr + r0 * w * 4, 4 * w, w, h, r0, r1 - r0, 2, (2, 1, 0))
            kernels.add(ctx.last_kernel())
        got = ctx.download(d_dst).reshape(h, 4 * w)
        assert np.array_equal(got, want), (kernels, np.argwhere(got != want)[:4])
    finally:
        ctx.set_variant("auto")


# ------------------------------------------------------------------ rgb2bayer
@pytest.mark.parametrize("w,h", [(9, 7), (16, 4), (130, 21), (1030, 9), (1024, 16), (3840, 6)])
def test_rgb2bayer_all_patterns(ctx, orc, rng, w, h):
    """gstrgb2bayer.c:254-267 (byte 3 / 1 / 2 of the ARGB pixel by row/column parity). Widths that are not a
    multiple of the 16 pixels a lane owns, rows that start off 16-byte alignment (w = 9: mosaic pitch 12),
    a 4K row, and a source pitch with padding."""
    for pad in (0, 12):
        src = frames.random_u8(rng, h, 4 * w + pad)
        mstride = frames.round_up_4(w)
        for fmt in BAYER_FORMATS:
            want = orc.rgb2bayer(src, w, h, fmt)
            d_src = ctx.upload(src)
            d_dst = ctx.alloc(h * mstride)
            ctx.rgb2bayer(d_src, src.shape[1], d_dst, mstride, w, h, BAYER_FORMATS[fmt])
            got = ctx.download(d_dst, h * mstride).reshape(h, mstride)
            assert np.array_equal(got[:, :w], want[:, :w]), (w, h, pad, fmt, np.argwhere(got[:, :w] != want[:, :w])[:4])


def test_rgb2bayer_batch_then_bayer2rgb_round_trip(ctx, orc, rng):
    """a batch of frames through rgb2bayer, and the size-independent property the plugin pair offers: mosaicing
    a demosaiced frame returns the original samples (each pixel keeps its own colour sample)"""
    w, h, n = 256, 48, 3
    mosaic = rng.integers(0, 256, (n, h, w), dtype=np.uint8)
    d_m = ctx.upload(mosaic)
    d_rgb = ctx.alloc(n * h * w * 4)
    ctx.bayer2rgb(d_m, w, d_rgb, 4 * w, w, h, BAYER_FORMATS["bggr"], RGB_OFFSETS["ARGB"], nframes=n)
    d_back = ctx.alloc(n * h * w)
    ctx.rgb2bayer(d_rgb, 4 * w, d_back, w, w, h, BAYER_FORMATS["bggr"], nframes=n)
    back = ctx.download(d_back, n * h * w).reshape(n, h, w)
    rgb = ctx.download(d_rgb, n * h * w * 4).reshape(n, h, 4 * w)
    for i in range(n):
        assert np.array_equal(back[i], orc.rgb2bayer(rgb[i], w, h, "bggr")[:, :w]), i
    assert np.array_equal(back, mosaic)
