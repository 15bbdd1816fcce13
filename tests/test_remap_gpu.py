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
