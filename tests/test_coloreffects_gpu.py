"""coloreffects / chromahold parity (bit-exact, in place), through the C-ABI."""
import numpy as np
import pytest

import frames
from oracle import PRESETS, RGB_OFFSETS

pytestmark = pytest.mark.gpu

FORMATS = ["RGB", "BGR", "BGRx", "xRGB", "ARGB", "RGBA", "xBGR", "ABGR", "RGBx", "BGRA", "AYUV"]


def run_ce(ctx, vf, frame, w, h, fmt, preset, nframes=1):
    d = ctx.upload(frame)
    stride = frame.shape[-1]
    if preset != "none":
        table, ml = vf.coloreffects_table(PRESETS[preset])
        if fmt == "AYUV":
            ctx.coloreffects_ayuv(d, w, h, stride, RGB_OFFSETS[fmt], table, ml, nframes=nframes)
        else:
            ps = 3 if fmt in ("RGB", "BGR") else 4
            ctx.coloreffects_rgb(d, w, h, stride, ps, RGB_OFFSETS[fmt], table, ml, nframes=nframes)
    return ctx.download(d).reshape(frame.shape)


@pytest.mark.parametrize("fmt", FORMATS)
@pytest.mark.parametrize("preset", list(PRESETS))
def test_presets_formats(ctx, vf, orc, rng, fmt, preset):
    ps = 3 if fmt in ("RGB", "BGR") else 4
    for (w, h) in [(33, 17), (64, 48), (1, 1), (5, 3)]:
        stride = frames.round_up_4(w * ps)        # RGB with w=33: non-zero row_wrap
        fr = frames.random_u8(rng, h, stride)
        got = run_ce(ctx, vf, fr, w, h, fmt, preset)
        want = orc.coloreffects(fr, w, h, fmt, preset)
        assert np.array_equal(got, want), (fmt, preset, w, h, ctx.last_kernel())


def test_tables_match_reference(vf, orc):
    import oracle
    gold = oracle.coloreffects_tables()
    for name, idx in PRESETS.items():
        if name == "none":
            continue
        t, ml = vf.coloreffects_table(idx)
        assert np.array_equal(t, gold[name]) and ml == oracle.PRESET_MAP_LUMA[name]


def test_all_yuv_values_ayuv(ctx, vf, orc):
    """every (Y,U,V) on a 64-step lattice + every Y: covers the matrix clamps"""
    y, u, v = np.meshgrid(np.arange(256), np.arange(0, 256, 5), np.arange(0, 256, 5), indexing="ij")
    px = np.stack([np.full(y.size, 200), y.reshape(-1), u.reshape(-1), v.reshape(-1)], 1).astype(np.uint8)
    w = 256
    h = px.shape[0] // w
    fr = px[: w * h].reshape(h, w * 4)
    for preset in ("sepia", "xpro", "yellowblue", "heat"):
        assert np.array_equal(run_ce(ctx, vf, fr, w, h, "AYUV", preset), orc.coloreffects(fr, w, h, "AYUV", preset))


def test_batch_contiguous_and_padded(ctx, vf, orc, rng):
    w, h, n = 64, 16, 3
    fr = rng.integers(0, 256, (n * h, w * 4), dtype=np.uint8)
    got = run_ce(ctx, vf, fr, w, h, "BGRx", "sepia", nframes=n).reshape(n, h, w * 4)
    for i in range(n):
        assert np.array_equal(got[i], orc.coloreffects(fr[i * h:(i + 1) * h], w, h, "BGRx", "sepia"))
    # padded rows (stride > 4*w): padding bytes must stay untouched
    frp = rng.integers(0, 256, (h, w * 4 + 16), dtype=np.uint8)
    got = run_ce(ctx, vf, frp, w, h, "RGBA", "xray")
    assert np.array_equal(got, orc.coloreffects(frp, w, h, "RGBA", "xray"))


@pytest.mark.parametrize("fmt", ["ARGB", "BGRA", "ABGR", "RGBA", "xRGB", "BGRx", "xBGR", "RGBx"])
def test_chromahold(ctx, orc, rng, fmt):
    for (w, h) in [(33, 17), (64, 48)]:
        fr = frames.random_u8(rng, h, 4 * w)
        fr[0, :64] = np.repeat(np.arange(16, dtype=np.uint8) * 16, 4)     # some greys (C == 0)
        for tgt in [(255, 0, 0), (128, 128, 128), (0, 200, 30), (10, 10, 250)]:
            for tol in (0, 30, 180):
                d = ctx.upload(fr)
                ctx.chromahold(d, w, h, 4 * w, RGB_OFFSETS[fmt], tgt, tol)
                got = ctx.download(d).reshape(fr.shape)
                assert np.array_equal(got, orc.chromahold(fr, w, h, fmt, tgt, tol)), (fmt, tgt, tol)


def test_chromahold_all_hues(ctx, orc):
    """all (r,g,b) on a lattice: every divisor C and every sign of the numerator"""
    r, g, b = np.meshgrid(np.arange(0, 256, 3), np.arange(0, 256, 3), np.arange(0, 256, 3), indexing="ij")
    px = np.stack([r.reshape(-1), g.reshape(-1), b.reshape(-1), np.full(r.size, 9)], 1).astype(np.uint8)
    w = 256
    h = px.shape[0] // w
    fr = px[: w * h].reshape(h, w * 4)
    for tol in (0, 17, 90):
        d = ctx.upload(fr)
        ctx.chromahold(d, w, h, 4 * w, RGB_OFFSETS["RGBA"], (40, 200, 90), tol)
        assert np.array_equal(ctx.download(d).reshape(fr.shape), orc.chromahold(fr, w, h, "RGBA", (40, 200, 90), tol))


@pytest.mark.parametrize("w,h,n", [(64, 64, 1), (640, 480, 1), (100, 41, 3), (1000, 33, 2)])
def test_tma_ring_and_grid_stride_paths(ctx, vf, orc, rng, w, h, n):
    """contiguous 4-byte frames of >= 16 KB stream through the TMA ring (stream.cuh); `direct` forces the
    grid-stride kernels: same bytes from both, in place"""
    fr = rng.integers(0, 256, (n * h, w * 4), dtype=np.uint8)
    for variant, suffix in (("auto", "_tma"), ("direct", "")):
        ctx.set_variant(variant)
        try:
            for fmt, preset in (("BGRx", "sepia"), ("RGBA", "xpro"), ("AYUV", "heat"), ("AYUV", "yellowblue")):
                got = run_ce(ctx, vf, fr, w, h, fmt, preset, nframes=n).reshape(n, h, w * 4)
                assert ctx.last_kernel().endswith("_tma") == (suffix == "_tma"), ctx.last_kernel()
                for i in range(n):
                    assert np.array_equal(got[i], orc.coloreffects(fr[i * h:(i + 1) * h], w, h, fmt, preset)), (variant, fmt, preset)
            d = ctx.upload(fr)
            ctx.chromahold(d, w, h, 4 * w, RGB_OFFSETS["xRGB"], (0, 200, 30), 30, nframes=n)
            assert ctx.last_kernel() == "chromahold" + suffix
            got = ctx.download(d).reshape(n, h, w * 4)
            for i in range(n):
                assert np.array_equal(got[i], orc.chromahold(fr[i * h:(i + 1) * h], w, h, "xRGB", (0, 200, 30), 30))
        finally:
            ctx.set_variant("auto")
