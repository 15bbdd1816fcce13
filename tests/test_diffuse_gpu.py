"""diffuse (gst/geometrictransform/gstdiffuse.c:151-231): the element that draws a fresh random displacement for every
pixel of every frame. The reference's draws (GLib's global generator, seeded by the OS) cannot be reproduced - parity is
unpinned by construction - so the tests pin what can be pinned:
  * given the draws, the frame is the reference's: the (angle, distance) of every pixel is recomputed on the host from
    the generator's definition, turned into the double map diffuse_map would have produced, and pushed through the
    reference's own do_map (oracle.remap): bit-exact for every policy, pixel stride and the AYUV fill;
  * the draws have the distribution the reference's have: angle uniform over 0..255, distance uniform in [0, 1),
    independent of each other, of the position and of the frame number."""
import math

import numpy as np
import pytest

import frames

gpu = pytest.mark.gpu

OFF = {"ignore": 0, "clamp": 1, "wrap": 2}
M64 = (1 << 64) - 1


def draws(seed, frame, npx):
    """numpy restatement of csrc/diffuse.cu's generator: splitmix64's finaliser over a Weyl sequence of the counter"""
    with np.errstate(over="ignore"):
        ctr = (np.uint64(frame) << np.uint64(32)) + np.arange(npx, dtype=np.uint64) + np.uint64(1)
        z = np.uint64(seed) + np.uint64(0x9E3779B97F4A7C15) * ctr
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    angle = (z >> np.uint64(56)).astype(np.int64)
    distance = ((z >> np.uint64(4)) & np.uint64((1 << 52) - 1)).astype(np.float64) * (1.0 / 4503599627370496.0)
    return angle, distance


def host_tables(scale):
    """diffuse_prepare (gstdiffuse.c:151-165) with this host's libm (math.sin, not numpy's vectorised one)"""
    s = np.array([scale * math.sin((math.pi * 2 * i) / 256.0) for i in range(256)])
    c = np.array([scale * math.cos((math.pi * 2 * i) / 256.0) for i in range(256)])
    return s, c


def model_map(seed, frame, w, h, scale):
    """the gdouble map diffuse_map (gstdiffuse.c:167-187) produces from these draws"""
    angle, distance = draws(seed, frame, w * h)
    s, c = host_tables(scale)
    x = np.tile(np.arange(w, dtype=np.float64), h)
    y = np.repeat(np.arange(h, dtype=np.float64), w)
    m = np.empty((h, w, 2), np.float64)
    m[..., 0] = (x + distance * s[angle]).reshape(h, w)
    m[..., 1] = (y + distance * c[angle]).reshape(h, w)
    return m


def test_generator_on_host_equals_its_numpy_restatement(vf):
    angle, distance = draws(0x1234, 7, 500)
    for px in (0, 1, 2, 17, 255, 499):
        a, d = vf.diffuse_draw(0x1234, 7, px)
        assert (a, d) == (int(angle[px]), float(distance[px]))
    s, c = vf.diffuse_tables(4.0)
    hs, hc = host_tables(4.0)
    assert np.array_equal(s, hs) and np.array_equal(c, hc)


@gpu
@pytest.mark.parametrize("ps", [1, 2, 3, 4])
@pytest.mark.parametrize("off_edge", ["ignore", "clamp", "wrap"])
def test_frame_is_the_references_do_map_of_the_drawn_map(ctx, vf, orc, rng, ps, off_edge):
    w, h, scale, seed = 100, 37, 4.0, 0xfeed
    stride = (w * ps + 3) // 4 * 4
    fr = frames.random_u8(rng, h, stride)
    s, c = vf.diffuse_tables(scale)
    src, dst = ctx.upload(fr), ctx.alloc(fr.size)
    for frame in (0, 5):
        ctx.diffuse(src, dst, w, h, ps, stride, s, c, OFF[off_edge], 0, seed, frame)
        got = ctx.download(dst, fr.size).reshape(h, stride)
        want = orc.remap(fr, model_map(seed, frame, w, h, scale), w, h, ps, off_edge, False)
        assert np.array_equal(got, want), (ps, off_edge, frame, np.argwhere(got != want)[:4])
    assert ctx.last_kernel() == "diffuse"


@gpu
def test_ayuv_fill_large_scale_and_batches(ctx, vf, orc, rng):
    """scale 40 on a 64x48 frame throws many pixels off the frame: ignore leaves the AYUV black (0xff108080 BE) there;
    a batch of frames draws with consecutive frame numbers"""
    w, h, scale, seed, n = 64, 48, 40.0, 99, 3
    fr = frames.random_u8(rng, n * h, 4 * w)
    s, c = vf.diffuse_tables(scale)
    src, dst = ctx.upload(fr), ctx.alloc(fr.size)
    ctx.diffuse(src, dst, w, h, 4, 4 * w, s, c, 0, 0x808010ff, seed, 10, nframes=n)
    got = ctx.download(dst, fr.size).reshape(n, h, 4 * w)
    filled = 0
    for i in range(n):
        want = orc.remap(fr[i * h:(i + 1) * h], model_map(seed, 10 + i, w, h, scale), w, h, 4, "ignore", True)
        assert np.array_equal(got[i], want), i
        filled += int((want.view(np.uint32) == 0x808010ff).sum())
    assert filled > n * w * h // 10


@gpu
def test_row_shards_draw_the_numbers_of_the_whole_frame(ctx, vf, rng):
    """a shard writes rows [r0, r1) drawing with GLOBAL pixel numbers: stacked shards == the whole frame"""
    w, h, seed = 96, 40, 5
    fr = frames.random_u8(rng, h, 4 * w)
    s, c = vf.diffuse_tables(6.0)
    src, whole = ctx.upload(fr), ctx.alloc(fr.size)
    ctx.diffuse(src, whole, w, h, 4, 4 * w, s, c, 1, 0, seed, 3)
    want = ctx.download(whole, fr.size).reshape(h, 4 * w)
    parts = []
    for r0, r1 in ((0, 13), (13, 14), (14, 40)):
        d = ctx.alloc((r1 - r0) * 4 * w)
        ctx.diffuse(src, d, w, r1 - r0, 4, 4 * w, s, c, 1, 0, seed, 3, first_row=r0, full_height=h)
        parts.append(ctx.download(d, (r1 - r0) * 4 * w).reshape(r1 - r0, 4 * w))
    assert np.array_equal(np.vstack(parts), want)


def test_draw_statistics(vf):
    """what g_random_int_range (0, 256) / g_random_double () promise: uniform angle, uniform distance in [0, 1),
    no dependence between them, on the pixel position or on the frame number"""
    n = 1 << 20
    a0, d0 = draws(1, 0, n)
    a1, d1 = draws(1, 1, n)
    assert d0.min() >= 0.0 and d0.max() < 1.0
    for a in (a0, a1):
        hist = np.bincount(a, minlength=256)
        chi2 = float(((hist - n / 256) ** 2 / (n / 256)).sum())
        assert chi2 < 340, chi2                                        # 255 degrees of freedom: P(chi2 > 340) ~ 3e-4
    for d in (d0, d1):
        assert abs(d.mean() - 0.5) < 0.002 and abs(d.var() - 1 / 12) < 0.001
        hist = np.bincount((d * 64).astype(np.int64), minlength=64)
        assert float(((hist - n / 64) ** 2 / (n / 64)).sum()) < 120      # 63 degrees of freedom
    corr = lambda u, v: abs(float(np.corrcoef(u, v)[0, 1]))
    assert corr(a0.astype(float), d0) < 0.005                            # angle vs distance
    assert corr(d0, d1) < 0.005 and corr(a0.astype(float), a1.astype(float)) < 0.005      # frame to frame
    assert corr(d0[:-1], d0[1:]) < 0.005 and corr(a0[:-1].astype(float), a0[1:].astype(float)) < 0.005   # neighbours
    assert corr(d0, np.arange(n, dtype=float)) < 0.005                   # position
    assert (a0 != a1).mean() > 0.99
    # another seed is another texture
    a2, _ = draws(2, 0, n)
    assert (a0 != a2).mean() > 0.99


@gpu
def test_diffuse_element(ctx, vf, orc, rng):
    """factory surface of gstdiffuse.c:213-231 (scale [1, G_MAXDOUBLE] 4, off-edge-pixels clamp), a new draw every
    frame, reproducible from (seed, frame); the reference's diffuse_prepare builds its tables once (:156-157), so a
    later change of `scale` never reaches them - reproduced"""
    w, h = 80, 50
    fr = frames.random_u8(rng, h, 4 * w)
    e = ctx.element("diffuse")
    assert e.get_property("scale") == 4.0 and int(e.get_property("off-edge-pixels")) == 1
    with pytest.raises(vf.B200vfError):
        e.set_property("scale", 0.5)
    e.set_caps("RGBA", "RGBA", w, h)
    seed, nxt = e.rng_state()
    assert nxt == 0
    f0 = e.transform(fr).reshape(h, 4 * w)
    f1 = e.transform(fr).reshape(h, 4 * w)
    assert e.rng_state() == (seed, 2) and not np.array_equal(f0, f1)
    assert np.array_equal(f0, orc.remap(fr, model_map(seed, 0, w, h, 4.0), w, h, 4, "clamp", False))
    assert np.array_equal(f1, orc.remap(fr, model_map(seed, 1, w, h, 4.0), w, h, 4, "clamp", False))
    # every output pixel is a source pixel at most `scale` away
    px_in = fr.reshape(h, w, 4).view(np.uint32)[..., 0]
    m = model_map(seed, 0, w, h, 4.0)
    assert np.abs(m[..., 0] - np.arange(w)[None, :]).max() < 4.0 and np.abs(m[..., 1] - np.arange(h)[:, None]).max() < 4.0
    assert np.isin(f0.reshape(h, w, 4).view(np.uint32)[..., 0], px_in).all()
    # rewind: the same frame again; wrap policy through the property
    e.set_rng_seed(0xabc, 7)
    e.set_property("off-edge-pixels", "wrap")
    e.set_property("scale", 9.0)                                         # too late for the tables
    g = e.transform(fr).reshape(h, 4 * w)
    assert np.array_equal(g, orc.remap(fr, model_map(0xabc, 7, w, h, 4.0), w, h, 4, "wrap", False))
    # an element negotiated AFTER the property change uses it
    e2 = ctx.element("diffuse")
    e2.set_property("scale", 9.0)
    e2.set_caps("RGBA", "RGBA", w, h)
    e2.set_rng_seed(0xabc, 7)
    g2 = e2.transform(fr).reshape(h, 4 * w)
    assert np.array_equal(g2, orc.remap(fr, model_map(0xabc, 7, w, h, 9.0), w, h, 4, "clamp", False))
    with pytest.raises(vf.B200vfError):
        ctx.element("fisheye").set_rng_seed(1)
