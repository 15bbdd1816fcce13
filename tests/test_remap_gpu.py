"""geometrictransform parity: host map builder vs the reference's map functions, and the CUDA gather
vs the reference's do_map loop (bit-exact integer copy)."""
import numpy as np
import pytest

import frames
import oracle as orc_mod

pytestmark = pytest.mark.gpu

OFF = {"ignore": 0, "clamp": 1, "wrap": 2}


def gpu_remap(ctx, vf, fr, index, w, h, ps, fill=0, nframes=1):
    d_src = ctx.upload(fr)
    d_dst = ctx.alloc(fr.size)
    d_idx = ctx.upload(index)
    ctx.remap(d_src, d_dst, d_idx, w, h, ps, fr.shape[-1], fill=fill, nframes=nframes)
    return ctx.download(d_dst, fr.size).reshape(fr.shape)


@pytest.mark.parametrize("w,h", [(64, 48), (100, 75)])
@pytest.mark.parametrize("ps", [1, 2, 3, 4])
@pytest.mark.parametrize("off_edge", ["ignore", "clamp", "wrap"])
def test_fisheye_all_pixel_strides(ctx, vf, orc, rng, w, h, ps, off_edge):
    m = vf.gt_build_map("fisheye", w, h)
    assert np.array_equal(m.view(np.uint64), orc.gt_map("fisheye", w, h).view(np.uint64))
    idx = vf.gt_resolve_map(m, w, h, OFF[off_edge])
    fr = frames.random_u8(rng, h, frames.round_up_4(w * ps))
    got = gpu_remap(ctx, vf, fr, idx, w, h, ps)
    want = orc.remap(fr, m, w, h, ps, off_edge, False)
    assert np.array_equal(got, want), (ctx.last_kernel(), np.argwhere(got != want)[:4])


def test_ayuv_fill(ctx, vf, orc, rng):
    w, h = 64, 48
    m = orc_mod.get("port").gt_map("fisheye", w, h) * 1.7 - 20.0        # pushes many pixels off the frame
    idx = vf.gt_resolve_map(m, w, h, 0)
    fr = frames.random_u8(rng, h, 4 * w)
    got = gpu_remap(ctx, vf, fr, idx, w, h, 4, fill=0x808010ff)
    assert np.array_equal(got, orc.remap(fr, m, w, h, 4, "ignore", True))
    assert (idx < 0).any()


def test_batch(ctx, vf, orc, rng):
    w, h, n = 64, 32, 3
    m = vf.gt_build_map("fisheye", w, h)
    idx = vf.gt_resolve_map(m, w, h, 1)
    fr = rng.integers(0, 256, (n * h, 4 * w), dtype=np.uint8)
    got = gpu_remap(ctx, vf, fr, idx, w, h, 4, nframes=n).reshape(n, h, 4 * w)
    for i in range(n):
        assert np.array_equal(got[i], orc.remap(fr[i * h:(i + 1) * h], m, w, h, 4, "clamp", False))


def test_fisheye_4k_index_and_frame(ctx, vf, orc, rng):
    w, h = 3840, 2160
    m = vf.gt_build_map("fisheye", w, h)
    assert np.array_equal(m.view(np.uint64), orc.gt_map("fisheye", w, h).view(np.uint64))
    idx = vf.gt_resolve_map(m, w, h, 1)
    fr = frames.random_u8(rng, h, 4 * w)
    got = gpu_remap(ctx, vf, fr, idx, w, h, 4)
    assert np.array_equal(got, orc.remap(fr, m, w, h, 4, "clamp", False))


def gpu_remap_packed(ctx, vf, fr, index, w, h, fill=0, nframes=1):
    packed, raw = vf.gt_pack_index(index, w, h)
    d_src = ctx.upload(fr)
    d_dst = ctx.alloc(fr.size)
    d_p = ctx.upload(packed)
    ctx.remap_packed(d_src, d_dst, d_p, w, h, fill=fill, nframes=nframes)
    assert ctx.last_kernel() == "remap4_packed"
    return ctx.download(d_dst, fr.size).reshape(fr.shape), raw


@pytest.mark.parametrize("w,h", [(300, 40), (128, 16), (100, 9), (257, 5), (5, 3)])
def test_packed_table_all_maps(ctx, vf, orc, rng, w, h):
    """the step-coded table drives the same gather: every map, every off-edge policy, ragged widths, AYUV fill"""
    import refprops
    fr = frames.random_u8(rng, h, 4 * w)
    for el, plist in refprops.CASES.items():
        if el == "diffuse":
            continue
        m = vf.gt_build_map(el, w, h, plist[0])
        for name, off in OFF.items():
            idx = vf.gt_resolve_map(m, w, h, off)
            got, raw = gpu_remap_packed(ctx, vf, fr, idx, w, h, fill=0x808010ff)
            want = orc.remap(fr, m, w, h, 4, name, True)
            assert np.array_equal(got, want), (el, name, raw, np.argwhere(got != want)[:4])


def test_packed_table_batch_and_raw_groups(ctx, vf, rng):
    w, h, n = 384, 24, 3
    idx = rng.integers(-1, w * h, (h, w), dtype=np.int32)                  # nothing codable: all groups raw
    smooth = vf.gt_resolve_map(vf.gt_build_map("twirl", w, h), w, h, 1)
    idx[::2] = smooth[::2]                                                 # every other row coded
    fr = rng.integers(0, 256, (n * h, 4 * w), dtype=np.uint8)
    got, raw = gpu_remap_packed(ctx, vf, fr, idx, w, h, fill=0x01020304, nframes=n)
    assert (w // 8) * (h // 2) <= raw < (w // 8) * h
    px = fr.reshape(n, h * w, 4)
    want = np.where((idx.reshape(-1) >= 0)[None, :, None], px[:, np.maximum(idx.reshape(-1), 0)], np.array([4, 3, 2, 1], np.uint8))
    assert np.array_equal(got.reshape(n, h * w, 4), want)


def test_packed_fisheye_4k_equals_plain(ctx, vf, rng):
    w, h = 3840, 2160
    idx = vf.gt_resolve_map(vf.gt_build_map("fisheye", w, h), w, h, 1)
    fr = frames.random_u8(rng, h, 4 * w)
    got, raw = gpu_remap_packed(ctx, vf, fr, idx, w, h)
    assert raw == 0
    assert np.array_equal(got, gpu_remap(ctx, vf, fr, idx, w, h, 4))


DEVICE_MAPS = ("mirror", "square", "stretch", "bulge", "tunnel", "perspective", "marble")
LIBM_MAPS = ("fisheye", "circle", "kaleidoscope", "pinch", "rotate", "sphere", "twirl", "waterripple")
MARBLE_CASES = [{}, {"x_scale": 7.0, "y_scale": 2.5, "turbulence": 3.0}]


def device_cases(element):
    import refprops
    return MARBLE_CASES if element == "marble" else refprops.CASES[element]


@pytest.mark.parametrize("element", DEVICE_MAPS)
@pytest.mark.parametrize("w,h", [(64, 48), (100, 75), (257, 131)])
def test_device_built_index_equals_host_table(ctx, vf, element, w, h):
    """the maps that need no libm are evaluated on the GPU in the reference's fp64 expression order;
    the int32 table must equal, entry for entry, the one resolved on the host from the (reference-bit-equal) double
    map - for every property set of the parity tests and every off-edge policy. (marble: libm-free once the host has
    built its lattice / displacement tables.)"""
    assert vf.gt_device_map_supported(element) == 1
    for props in device_cases(element):
        m = vf.gt_build_map(element, w, h, props)
        for name, off in OFF.items():
            want = vf.gt_resolve_map(m, w, h, off)
            d = vf.gt_build_index_device(ctx, element, w, h, props, off)
            got = ctx.download(d, w * h * 4, dtype=np.int32).reshape(h, w)
            assert np.array_equal(got, want.reshape(h, w)), (element, props, name, np.argwhere(got != want.reshape(h, w))[:5])
    assert ctx.last_kernel() == "gt_index_device"


def test_device_built_index_at_4k_and_8k(ctx, vf):
    for (w, h) in [(3840, 2160), (7680, 4320)]:
        for element, props in (("bulge", {"zoom": 7.5, "x_center": 0.3, "radius": 0.6}), ("square", {}), ("tunnel", {}), ("marble", {})):
            want = vf.gt_resolve_map(vf.gt_build_map(element, w, h, props), w, h, 1)
            got = ctx.download(vf.gt_build_index_device(ctx, element, w, h, props, 1), w * h * 4, dtype=np.int32)
            assert np.array_equal(got, want.reshape(-1)), (element, w)


def build_certified(ctx, vf, element, w, h, props, off):
    """the device table, or None when the build reports that it cannot certify the map"""
    try:
        d = vf.gt_build_index_device(ctx, element, w, h, props, off)
    except vf.B200vfError as err:
        assert err.status == vf.E_UNSUPPORTED, err
        return None
    return ctx.download(d, w * h * 4, dtype=np.int32).reshape(h, w)


@pytest.mark.parametrize("element", LIBM_MAPS)
@pytest.mark.parametrize("w,h", [(64, 48), (100, 75), (257, 131)])
def test_certified_libm_tables_equal_host_tables(ctx, vf, element, w, h):
    """The maps that call libm: CUDA's pow / atan2 / sin ... differ from glibc's in the last ulps and do_map truncates, so
    the kernel certifies each entry (coordinate further from an integer than a bound on the disagreement) and takes the
    others from the host's map function. The patched table must equal the host's, entry for entry; the only property
    set the build may refuse is rotate at angle 0, where EVERY coordinate sits within 1e-13 of an integer."""
    import refprops
    assert vf.gt_device_map_supported(element) == 2
    for props in refprops.CASES[element]:
        m = vf.gt_build_map(element, w, h, props)
        for name, off in OFF.items():
            want = vf.gt_resolve_map(m, w, h, off).reshape(h, w)
            got = build_certified(ctx, vf, element, w, h, props, off)
            if got is None:
                assert element == "rotate" and not props, (element, props)
                continue
            assert vf.gt_device_last_uncertain() <= max(4096, w * h // 64)
            assert np.array_equal(got, want), (element, props, name, np.argwhere(got != want)[:5])


def test_certified_libm_tables_over_random_properties(ctx, vf):
    """a seeded sweep over the property ranges of the libm maps (incl. refraction < 1, where sphere's asin leaves its
    domain, negative pinch intensities, the angle wrap of circle) at a size whose centre is not a pixel"""
    import refprops
    rng = np.random.default_rng(7)
    w, h = 321, 203
    ranges = {
        "circle": {"angle": (-3.2, 3.2), "spread_angle": (0.2, 6.3), "height": (1, 80), "x_center": (0, 1), "y_center": (0, 1), "radius": (0, 1)},
        "kaleidoscope": {"angle": (-3.2, 3.2), "angle2": (-3.2, 3.2), "sides": (2, 9), "x_center": (0, 1), "radius": (0, 1)},
        "pinch": {"intensity": (-1, 1), "x_center": (0, 1), "y_center": (0, 1), "radius": (0.05, 1)},
        "rotate": {"angle": (-6.3, 6.3)},
        "sphere": {"refraction": (0.3, 3.0), "x_center": (0, 1), "radius": (0.05, 1)},
        "twirl": {"angle": (-6.3, 6.3), "y_center": (0, 1), "radius": (0.05, 1)},
        "waterripple": {"amplitude": (-40, 40), "phase": (-7, 7), "wavelength": (2, 60), "x_center": (0, 1), "radius": (0.05, 1)},
    }
    refused = 0
    for element, rr in ranges.items():
        for trial in range(6):
            props = {k: (float(np.floor(rng.uniform(*v))) if k in ("sides", "height") else float(rng.uniform(*v))) for k, v in rr.items()}
            off = int(rng.integers(0, 3))
            want = vf.gt_resolve_map(vf.gt_build_map(element, w, h, props), w, h, off).reshape(h, w)
            got = build_certified(ctx, vf, element, w, h, props, off)
            if got is None:
                refused += 1
                continue
            assert np.array_equal(got, want), (element, props, off, np.argwhere(got != want)[:5])
    assert refused <= 4, refused


def test_certified_libm_tables_at_4k_and_8k(ctx, vf):
    """fisheye (BASELINE configs[3]) and three more at the BASELINE sizes; how many entries the host had to supply"""
    for (w, h) in [(3840, 2160), (7680, 4320)]:
        for element, props in (("fisheye", {}), ("twirl", {"angle": -2.0}), ("sphere", {}), ("kaleidoscope", {"sides": 7, "x_center": 0.25})):
            want = vf.gt_resolve_map(vf.gt_build_map(element, w, h, props), w, h, 1)
            got = ctx.download(vf.gt_build_index_device(ctx, element, w, h, props, 1), w * h * 4, dtype=np.int32)
            assert np.array_equal(got, want.reshape(-1)), (element, w)
            assert vf.gt_device_last_uncertain() < w * h // 256, (element, vf.gt_device_last_uncertain())


def test_uncertifiable_map_falls_back_to_the_host_table(ctx, vf, orc, rng):
    """rotate at its default angle 0: the build refuses, the element builds the table on the host, the frame is the reference's"""
    w, h = 160, 120
    with pytest.raises(vf.B200vfError) as err:
        vf.gt_build_index_device(ctx, "rotate", w, h)
    assert err.value.status == vf.E_UNSUPPORTED and vf.gt_device_last_uncertain() > w * h // 64
    with pytest.raises(vf.B200vfError):
        vf.gt_build_index_device(ctx, "diffuse", w, h)
    assert vf.gt_device_map_supported("diffuse") == 0
    fr = frames.random_u8(rng, h, 4 * w)
    e = ctx.element("rotate")
    e.set_caps("RGBA", "RGBA", w, h)
    got = e.transform(fr).reshape(h, 4 * w)
    policy = {v: k for k, v in OFF.items()}[int(e.get_property("off-edge-pixels"))]
    assert np.array_equal(got, orc.remap(fr, vf.gt_build_map("rotate", w, h), w, h, 4, policy, False))


def test_element_rebuilds_its_table_on_the_gpu_when_a_property_moves(ctx, vf, orc, rng, monkeypatch):
    """a GstController tick = set_property + next frame: the bulge element's table is rebuilt by one kernel
    (no host map, no upload) and the frame equals the reference's; the host route gives the same bytes"""
    w, h = 320, 200
    fr = frames.random_u8(rng, h, 4 * w)
    e = ctx.element("bulge")
    e.set_caps("RGBA", "RGBA", w, h)
    outs = []
    for zoom in (3.0, 4.5, 9.0):
        e.set_property("zoom", zoom)
        l0 = ctx.launch_count()
        outs.append(e.transform(fr).reshape(h, 4 * w))
        assert ctx.launch_count() - l0 == 2                       # gt_index_device + remap4
        want = orc.remap(fr, vf.gt_build_map("bulge", w, h, {"zoom": zoom}), w, h, 4, "clamp", False)
        assert np.array_equal(outs[-1], want), zoom
    monkeypatch.setenv("B200VF_GT_HOST_MAPS", "1")
    e2 = ctx.element("bulge")
    e2.set_caps("RGBA", "RGBA", w, h)
    e2.set_property("zoom", 9.0)
    assert np.array_equal(e2.transform(fr).reshape(h, 4 * w), outs[-1])
