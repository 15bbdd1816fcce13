"""gaudieffects point ops parity (bit-exact), through the C-ABI."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def all_bytes_px():
    b = np.arange(256, dtype=np.uint32)
    return (b | (b << 8) | (b << 16) | (b << 24))


def run_lut(ctx, src_u32, lut):
    src = np.ascontiguousarray(src_u32, np.uint32)
    d_src = ctx.upload(src)
    d_dst = ctx.alloc(src.nbytes)
    ctx.lut4(d_src, d_dst, src.size, lut)
    return ctx.download(d_dst, dtype=np.uint32).reshape(src.shape)


@pytest.mark.parametrize("n", [1, 3, 4, 5, 255, 256, 4099, 64 * 48, 1 << 20])
def test_burn_sizes(ctx, orc, vf, rng, n):
    src = rng.integers(0, 2 ** 32, n, dtype=np.uint32)
    got = run_lut(ctx, src, vf.lut_burn(175))
    assert np.array_equal(got, orc.burn(src, 175))


@pytest.mark.parametrize("adj", [0, 1, 175, 255, 256])
def test_burn_adjustment(ctx, orc, vf, rng, adj):
    src = np.concatenate([all_bytes_px(), rng.integers(0, 2 ** 32, 5000, dtype=np.uint32)])
    assert np.array_equal(run_lut(ctx, src, vf.lut_burn(adj)), orc.burn(src, adj))


def test_dodge(ctx, orc, vf, rng):
    src = np.concatenate([all_bytes_px(), rng.integers(0, 2 ** 32, 5000, dtype=np.uint32)])
    assert np.array_equal(run_lut(ctx, src, vf.lut_dodge()), orc.dodge(src))


@pytest.mark.parametrize("a,b", [(200, 1), (0, 0), (256, 256), (37, 255)])
def test_chromium(ctx, orc, vf, rng, a, b):
    src = np.concatenate([all_bytes_px(), rng.integers(0, 2 ** 32, 5000, dtype=np.uint32)])
    assert np.array_equal(run_lut(ctx, src, vf.lut_chromium(a, b)), orc.chromium(src, a, b))


@pytest.mark.parametrize("t,s,e", [(127, 50, 185), (50, 50, 185), (185, 50, 185), (100, 100, 100), (127, 185, 50),
                                   (0, 0, 256), (256, 0, 256), (10, 200, 30)])
def test_solarize(ctx, orc, vf, rng, t, s, e):
    src = np.concatenate([all_bytes_px(), rng.integers(0, 2 ** 32, 5000, dtype=np.uint32)])
    assert np.array_equal(run_lut(ctx, src, vf.lut_solarize(t, s, e)), orc.solarize(src, t, s, e))


def test_misaligned_buffers(ctx, orc, vf, rng):
    """src and dst not congruent mod 16 -> scalar path; still bit-exact."""
    n = 1000
    src = rng.integers(0, 2 ** 32, n + 8, dtype=np.uint32)
    d_src = ctx.upload(src)
    d_dst = ctx.alloc(4 * (n + 8))
    ctx.lut4(d_src.ptr + 4, d_dst.ptr + 8, n, vf.lut_burn(175))
    got = ctx.download(d_dst, dtype=np.uint32)[2:2 + n]
    assert np.array_equal(got, orc.burn(src[1:1 + n], 175))
    ctx.lut4(d_src.ptr + 4, d_dst.ptr + 4, n, vf.lut_burn(175))     # congruent but not 16-aligned: head peel
    got = ctx.download(d_dst, dtype=np.uint32)[1:1 + n]
    assert np.array_equal(got, orc.burn(src[1:1 + n], 175))


@pytest.mark.parametrize("factor", [1, 2, 100, 175])
def test_exclusion(ctx, orc, rng, factor):
    src = np.concatenate([all_bytes_px(), rng.integers(0, 2 ** 32, 64 * 48 + 3, dtype=np.uint32)])
    d_src = ctx.upload(src)
    d_dst = ctx.alloc(src.nbytes)
    ctx.exclusion(d_src, d_dst, src.size, factor)
    assert np.array_equal(ctx.download(d_dst, dtype=np.uint32), orc.exclusion(src, factor))


def test_exclusion_all_red_green_pairs(ctx, orc):
    """the cross term (green*red)/factor for every (r,g) pair and every factor class"""
    r, g = np.meshgrid(np.arange(256, dtype=np.uint32), np.arange(256, dtype=np.uint32))
    src = ((r << 16) | (g << 8) | (r ^ g)).reshape(-1)
    d_src = ctx.upload(src)
    d_dst = ctx.alloc(src.nbytes)
    for factor in [1, 2, 3, 7, 64, 127, 128, 174, 175]:
        ctx.exclusion(d_src, d_dst, src.size, factor)
        assert np.array_equal(ctx.download(d_dst, dtype=np.uint32), orc.exclusion(src, factor)), factor


@pytest.mark.parametrize("w,h", [(1, 1), (1, 7), (7, 1), (4, 4), (64, 48), (130, 33), (256, 40), (640, 480)])
@pytest.mark.parametrize("erode", [False, True])
def test_dilate(ctx, orc, rng, w, h, erode):
    src = rng.integers(0, 2 ** 32, (h, w), dtype=np.uint32)
    if w >= 64:
        src[::3, ::5] = src[1, 1]          # ties: the strict compare keeps the earlier candidate
    d_src = ctx.upload(src)
    d_dst = ctx.alloc(src.nbytes)
    ctx.dilate(d_src, d_dst, w, h, erode)
    got = ctx.download(d_dst, dtype=np.uint32).reshape(h, w)
    assert np.array_equal(got, orc.dilate(src, erode)), ctx.last_kernel()


@pytest.mark.parametrize("w,h", [(4, 2), (128, 32), (132, 33), (260, 70), (1024, 100), (3840, 64)])
def test_dilate_tma_and_direct_paths(ctx, orc, rng, w, h):
    """widths that are multiples of 4 take the TMA-fed kernel; `direct` forces the register-marching one: both exact"""
    n = 2
    src = rng.integers(0, 2 ** 32, (n, h, w), dtype=np.uint32)
    src[:, ::3, ::5] = src[0, 0, 0]
    d_src = ctx.upload(src)
    d_dst = ctx.alloc(src.nbytes)
    for erode in (False, True):
        want = np.stack([orc.dilate(src[i], erode) for i in range(n)])
        for variant, kernel in (("auto", "dilate_tma"), ("direct", "dilate")):
            ctx.set_variant(variant)
            try:
                ctx.dilate(d_src, d_dst, w, h, erode, nframes=n)
            finally:
                ctx.set_variant("auto")
            assert ctx.last_kernel() == kernel
            got = ctx.download(d_dst, dtype=np.uint32).reshape(n, h, w)
            assert np.array_equal(got, want), (variant, erode, np.argwhere(got != want)[:4])


def test_dilate_batch_and_shard_below(ctx, orc, rng):
    w, h, n = 128, 32, 3
    src = rng.integers(0, 2 ** 32, (n, h, w), dtype=np.uint32)
    d_src = ctx.upload(src)
    d_dst = ctx.alloc(src.nbytes)
    ctx.dilate(d_src, d_dst, w, h, False, nframes=n)
    got = ctx.download(d_dst, dtype=np.uint32).reshape(n, h, w)
    for i in range(n):
        assert np.array_equal(got[i], orc.dilate(src[i], False))
    # a shard: the top 20 rows, with row 20 supplied as the row below
    full = orc.dilate(src[0], False)
    d_below = ctx.upload(src[0, 20])
    ctx.dilate(d_src, d_dst, w, 20, False, nframes=1, below=d_below)
    got = ctx.download(d_dst, dtype=np.uint32)[: 20 * w].reshape(20, w)
    assert np.array_equal(got, full[:20])


@pytest.mark.parametrize("n", [4096, 4099, 70001, (1 << 20) + 7])
def test_point_ops_tma_ring_and_grid_stride_paths(ctx, orc, vf, rng, n):
    """from 1024 16-byte groups on, lut4 and exclusion stream through the TMA ring (stream.cuh);
    `direct` forces the grid-stride kernels. Both bit-exact, ragged sizes included."""
    src = rng.integers(0, 2 ** 32, n, dtype=np.uint32)
    d_src = ctx.upload(src)
    d_dst = ctx.alloc(src.nbytes)
    for variant, k_lut, k_ex in (("auto", "lut4_t", "exclusion_t"), ("direct", "lut4", "exclusion")):
        ctx.set_variant(variant)
        try:
            ctx.lut4(d_src, d_dst, n, vf.lut_burn(175))
            assert ctx.last_kernel().startswith(k_lut), ctx.last_kernel()
            got = ctx.download(d_dst, dtype=np.uint32)
            assert np.array_equal(got, orc.burn(src, 175)), variant
            ctx.exclusion(d_src, d_dst, n, 100)
            assert ctx.last_kernel().startswith(k_ex), ctx.last_kernel()
            got = ctx.download(d_dst, dtype=np.uint32)
            assert np.array_equal(got, orc.exclusion(src, 100)), variant
        finally:
            ctx.set_variant("auto")
    # in place (dst == src) through the ring: a chunk is written back only after it was copied out
    ctx.lut4(d_src, d_src, n, vf.lut_burn(175))
    assert np.array_equal(ctx.download(d_src, dtype=np.uint32), orc.burn(src, 175))
