"""The element mirror: factory names, properties, caps and the transform vfunc on host buffers,
written the way the reference's GstHarness tests would read (push a buffer, compare bytes)."""
import numpy as np
import pytest

import frames
from oracle import PRESETS

pytestmark = pytest.mark.gpu


def test_bayer2rgb_element_host_path(ctx, orc, rng):
    e = ctx.element("bayer2rgb")
    w, h, n = 640, 480, 4
    e.set_caps("bggr", "BGRx", w, h)
    assert e.unit_size() == (w * h, w * h * 4)
    src = rng.integers(0, 256, (n, h, w), dtype=np.uint8)
    out = e.transform(src, n).reshape(n, h, 4 * w)
    for i in range(n):
        assert np.array_equal(out[i], orc.bayer2rgb(src[i], w, h, "bggr", "BGRx"))


def test_bayer2rgb_odd_pitch_width(ctx, orc, rng):
    e = ctx.element("bayer2rgb")
    w, h = 130, 20                                   # pitch ROUND_UP_4(130) = 132: the direct kernel
    e.set_caps("grbg", "ARGB", w, h)
    src = frames.random_u8(rng, h, 132)
    out = e.transform(src).reshape(h, 4 * w)
    assert np.array_equal(out, orc.bayer2rgb(src, w, h, "grbg", "ARGB"))


@pytest.mark.parametrize("name,props,ref", [
    ("burn", {"adjustment": 90}, lambda o, s: o.burn(s, 90)),
    ("dodge", {}, lambda o, s: o.dodge(s)),
    ("chromium", {"edge-a": 37, "edge-b": 255}, lambda o, s: o.chromium(s, 37, 255)),
    ("solarize", {"threshold": 10, "start": 200, "end": 30}, lambda o, s: o.solarize(s, 10, 200, 30)),
    ("exclusion", {"factor": 100}, lambda o, s: o.exclusion(s, 100)),
])
def test_gaudi_point_elements(ctx, orc, rng, name, props, ref):
    e = ctx.element(name)
    for k, v in props.items():
        e.set_property(k, v)
    w, h = 64, 48
    e.set_caps("BGRx", "BGRx", w, h)
    src = rng.integers(0, 2 ** 32, (h, w), dtype=np.uint32)
    out = e.transform(src.view(np.uint8)).view(np.uint32).reshape(h, w)
    assert np.array_equal(out, ref(orc, src))


def test_defaults_are_the_reference_defaults(ctx, orc, rng):
    w, h = 64, 48
    src = rng.integers(0, 2 ** 32, (h, w), dtype=np.uint32)
    for name, ref in [("burn", orc.burn(src)), ("chromium", orc.chromium(src)), ("solarize", orc.solarize(src)),
                      ("exclusion", orc.exclusion(src)), ("dilate", orc.dilate(src, False))]:
        e = ctx.element(name)
        e.set_caps("RGBx", "RGBx", w, h)
        out = e.transform(src.view(np.uint8)).view(np.uint32).reshape(h, w)
        assert np.array_equal(out, ref), name


def test_dilate_erode_property(ctx, orc, rng):
    e = ctx.element("dilate")
    e.set_property("erode", 1)
    w, h = 64, 48
    e.set_caps("BGRx", "BGRx", w, h)
    src = rng.integers(0, 2 ** 32, (h, w), dtype=np.uint32)
    out = e.transform(src.view(np.uint8)).view(np.uint32).reshape(h, w)
    assert np.array_equal(out, orc.dilate(src, True))


def test_gaussianblur_element_is_ayuv_only(ctx, vf, orc, rng):
    e = ctx.element("gaussianblur")
    with pytest.raises(vf.B200vfError) as err:
        e.set_caps("BGRx", "BGRx", 64, 48)           # the reference's pad template is AYUV only
    assert err.value.status == vf.E_UNSUPPORTED
    w, h = 64, 48
    e.set_caps("AYUV", "AYUV", w, h)
    e.set_property("sigma", 2.5)
    src = frames.random_u8(rng, h, 4 * w)
    out = e.transform(src).reshape(h, 4 * w)
    assert np.array_equal(out, orc.gaussblur(src, w, h, 2.5, 1))
    e.set_property("sigma", 0.0)                     # sigma 0: the element only copies
    assert np.array_equal(e.transform(src).reshape(h, 4 * w), src)


@pytest.mark.parametrize("fmt", ["RGB", "BGRx", "ARGB", "AYUV"])
@pytest.mark.parametrize("preset", ["none", "sepia", "xpro"])
def test_coloreffects_element(ctx, orc, rng, fmt, preset):
    e = ctx.element("coloreffects")
    e.set_property("preset", preset)
    assert e.get_property("preset") == PRESETS[preset]
    w, h = 33, 17
    e.set_caps(fmt, fmt, w, h)
    ps = 3 if fmt == "RGB" else 4
    src = frames.random_u8(rng, h, frames.round_up_4(w * ps))
    assert e.unit_size()[0] == src.size
    out = e.transform(src).reshape(src.shape)
    assert np.array_equal(out, orc.coloreffects(src, w, h, fmt, preset))


def test_chromahold_element(ctx, orc, rng):
    e = ctx.element("chromahold")
    for k, v in {"target-r": 0, "target-g": 200, "target-b": 30, "tolerance": 17}.items():
        e.set_property(k, v)
    w, h = 64, 48
    e.set_caps("xRGB", "xRGB", w, h)
    src = frames.random_u8(rng, h, 4 * w)
    assert np.array_equal(e.transform(src).reshape(src.shape), orc.chromahold(src, w, h, "xRGB", (0, 200, 30), 17))


def test_fisheye_element_default_is_clamp(ctx, orc, rng):
    e = ctx.element("fisheye")
    assert e.get_property("off-edge-pixels") == 1
    w, h = 100, 75
    e.set_caps("RGBA", "RGBA", w, h)
    src = frames.random_u8(rng, h, 4 * w)
    m = orc.gt_map("fisheye", w, h)
    assert np.array_equal(e.transform(src).reshape(src.shape), orc.remap(src, m, w, h, 4, "clamp", False))
    e.set_property("off-edge-pixels", "wrap")        # property change -> needs_remap
    assert np.array_equal(e.transform(src).reshape(src.shape), orc.remap(src, m, w, h, 4, "wrap", False))
    e.set_caps("GRAY8", "GRAY8", w, h)
    g = frames.random_u8(rng, h, w)
    assert np.array_equal(e.transform(g).reshape(g.shape), orc.remap(g, m, w, h, 1, "wrap", False))


def test_property_and_caps_errors(ctx, vf):
    e = ctx.element("burn")
    with pytest.raises(vf.B200vfError) as err:
        e.set_property("adjustment", 257)
    assert err.value.status == vf.E_PROPERTY
    with pytest.raises(vf.B200vfError):
        e.set_property("no-such-property", 1)
    with pytest.raises(vf.B200vfError) as err:
        e.transform(np.zeros(16, np.uint8))          # GST_FLOW_NOT_NEGOTIATED analogue
    assert err.value.status == vf.E_NOT_NEGOTIATED
    with pytest.raises(vf.B200vfError) as err:
        e.set_caps("AYUV", "AYUV", 8, 8)
    assert err.value.status == vf.E_UNSUPPORTED
    with pytest.raises(vf.B200vfError) as err:
        ctx.element("no-such-element")
    assert err.value.status == vf.E_UNSUPPORTED


def test_chain_bayer_coloreffects_solarize_fused_equals_unfused(ctx, vf, orc, rng):
    """BASELINE.json configs[4] at a test size: the fused kernel == the three elements run one by one
    == the oracle chain, for a luma preset and a per-channel preset."""
    w, h = 512, 96
    src = frames.random_u8(rng, h, w)
    rgb = orc.bayer2rgb(src, w, h, "bggr", "BGRx")
    for preset in ("sepia", "xpro"):
        want = orc.solarize(orc.coloreffects(rgb, w, h, "BGRx", preset).view(np.uint32)).view(np.uint8).reshape(h, 4 * w)
        table, ml = vf.coloreffects_table(PRESETS[preset])
        sol = vf.lut_solarize()
        d_src = ctx.upload(src)
        d_dst = ctx.alloc(h * w * 4)
        if ml:
            ctx.bayer2rgb_fused(d_src, w, d_dst, 4 * w, w, h, 0, (2, 1, 0), luma_table=table, lut=sol)
        else:
            ce = np.zeros((4, 256), np.uint8)
            ce[2], ce[1], ce[0], ce[3] = table[0::3], table[1::3], table[2::3], np.arange(256)   # BGRx: R at byte 2
            ctx.bayer2rgb_fused(d_src, w, d_dst, 4 * w, w, h, 0, (2, 1, 0), lut=vf.lut_compose(ce, sol))
        got = ctx.download(d_dst).reshape(h, 4 * w)
        assert np.array_equal(got, want), (preset, ctx.last_kernel())
        # unfused: three elements
        e1, e2, e3 = ctx.element("bayer2rgb"), ctx.element("coloreffects"), ctx.element("solarize")
        e1.set_caps("bggr", "BGRx", w, h); e2.set_caps("BGRx", "BGRx", w, h); e3.set_caps("BGRx", "BGRx", w, h)
        e2.set_property("preset", preset)
        out = e3.transform(e2.transform(e1.transform(src)))
        assert np.array_equal(out.reshape(h, 4 * w), want), preset


def test_frames_with_foreign_strides_and_plane_offsets(ctx, vf, orc, rng):
    """GstVideoMeta layouts (padded strides, moved planes): the reference elements honour GST_VIDEO_FRAME_PLANE_STRIDE /
    _PLANE_DATA (gstcoloreffects.c:315-329, gstzebrastripe.c:219-243); transform_host_layout repacks into the default
    layout in HBM and back. Same bytes as the default-layout call, padding untouched."""
    w, h = 100, 37
    # packed, in place: coloreffects on RGB (pstride 3) with rows padded to 512 bytes
    e = ctx.element("coloreffects")
    e.set_caps("RGB", "RGB", w, h)
    e.set_property("preset", "sepia")
    lay = e.default_layout(0)
    assert (lay.n_planes, lay.stride[0], lay.row_bytes[0], lay.rows[0]) == (1, 300, 300, h)
    fr = frames.random_u8(rng, h, 300)
    want = e.transform(fr.copy()).reshape(h, 300)
    big = rng.integers(0, 256, (h, 512), dtype=np.uint8)
    big[:, :300] = fr
    before = big.copy()
    lay.stride[0] = 512
    e.transform_layout(big, lay, big, lay)
    assert np.array_equal(big[:, :300], want) and np.array_equal(big[:, 300:], before[:, 300:])
    # planar, in place: zebrastripe on I420 with every plane moved and re-strided
    z = ctx.element("zebrastripe")
    z.set_caps("I420", "I420", w, h)
    d = z.default_layout(0)
    assert d.n_planes == 3 and d.stride[0] == 100 and d.stride[1] == 52 and d.offset[1] == 100 * 38
    size = d.offset[2] + d.stride[2] * d.rows[2]
    packed = rng.integers(0, 256, size, dtype=np.uint8)
    z2 = ctx.element("zebrastripe")
    z2.set_caps("I420", "I420", w, h)
    want = z.transform(packed.copy())
    moved = vf.FrameLayout()
    moved.n_planes = 3
    buf = rng.integers(0, 256, 40000, dtype=np.uint8)
    pos = 64
    for i in range(3):
        moved.offset[i], moved.stride[i] = pos, d.stride[i] + 28
        for r in range(d.rows[i]):
            buf[pos + r * moved.stride[i]: pos + r * moved.stride[i] + d.row_bytes[i]] = packed[d.offset[i] + r * d.stride[i]: d.offset[i] + r * d.stride[i] + d.row_bytes[i]]
        pos += moved.stride[i] * d.rows[i] + 128
    z2.transform_layout(buf, moved, buf, moved)
    for i in range(3):
        for r in range(d.rows[i]):
            got = buf[moved.offset[i] + r * moved.stride[i]: moved.offset[i] + r * moved.stride[i] + d.row_bytes[i]]
            assert np.array_equal(got, want[d.offset[i] + r * d.stride[i]: d.offset[i] + r * d.stride[i] + d.row_bytes[i]]), (i, r)
    # out of place with different layouts on both sides: bayer2rgb, mosaic rows padded to 256, RGBA rows to 1024
    b = ctx.element("bayer2rgb")
    b.set_caps("grbg", "RGBA", w, h)
    src = frames.random_u8(rng, h, 100)
    want = orc.bayer2rgb(src, w, h, "grbg", "RGBA")
    sin = np.zeros((h, 256), np.uint8); sin[:, :100] = src
    sout = np.full((h, 1024), 7, np.uint8)
    li, lo = b.default_layout(0), b.default_layout(1)
    li.stride[0], lo.stride[0] = 256, 1024
    b.transform_layout(sin, li, sout, lo)
    assert np.array_equal(sout[:, :400], want) and (sout[:, 400:] == 7).all()
