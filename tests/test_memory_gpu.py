"""b200vf_memory - the device-memory object the elements share (modelled on GstCudaMemory,
sys/nvcodec/gstcudamemory.c:257-407) - and the element transform on memories: frames stay in HBM between elements,
per-pixel elements join pending chains that are launched as one fused kernel (BASELINE.json configs[4])."""
import numpy as np
import pytest

import frames

pytestmark = pytest.mark.gpu

PRESETS = {"heat": 1, "sepia": 2, "xray": 3, "xpro": 4, "yellowblue": 5}


def counts_delta(ctx, before):
    now = ctx.transfer_counts()
    return tuple(a - b for a, b in zip(now, before))


def test_transfer_flags_follow_gstcudamemory(ctx, vf, rng):
    """host WRITE map -> NEED_UPLOAD at unmap; device READ map uploads once; device WRITE map -> NEED_DOWNLOAD; host READ
    map downloads once; a second host read costs nothing (cuda_mem_map / cuda_mem_unmap_full, gstcudamemory.c:331-407)"""
    n = 1 << 20
    data = rng.integers(0, 256, n, dtype=np.uint8)
    m = ctx.memory(n)
    c0 = ctx.transfer_counts()
    m.write(data)
    assert m.flags == vf.NEED_UPLOAD and counts_delta(ctx, c0) == (0, 0, 0, 0)        # nothing moved yet
    d = m.map(vf.MAP_READ | vf.MAP_DEVICE)
    m.unmap()
    assert d and m.flags == 0 and counts_delta(ctx, c0) == (1, n, 0, 0)
    assert np.array_equal(m.read(), data) and counts_delta(ctx, c0) == (1, n, 0, 0)   # both copies current: no download
    d2 = m.map(vf.MAP_WRITE | vf.MAP_DEVICE)
    m.unmap()
    assert d2 == d and m.flags == vf.NEED_DOWNLOAD
    ctx.synchronize()
    assert np.array_equal(m.read(), data) and counts_delta(ctx, c0) == (1, n, 1, n)   # (the device bytes were not changed)
    assert np.array_equal(m.read(), data) and counts_delta(ctx, c0) == (1, n, 1, n)
    m.close()


def test_pool_hands_out_memories_and_takes_them_back(ctx, vf):
    pool = ctx.pool(4096, 3)
    mems = [pool.acquire_memory() for _ in range(3)]
    assert len({m.map(vf.MAP_READ | vf.MAP_DEVICE) for m in mems}) == 3
    for m in mems:
        m.unmap()
    with pytest.raises(vf.B200vfError) as err:
        pool.acquire_memory()
    assert err.value.status == vf.E_NOMEM
    mems[1].close()
    again = pool.acquire_memory()
    assert again.nbytes == 4096
    pool.close()


@pytest.mark.parametrize("preset", ["sepia", "xpro"])
@pytest.mark.parametrize("size", [(512, 96), (7680, 4320)])
def test_chain_through_memories_is_one_upload_one_launch_one_download(ctx, vf, orc, rng, preset, size):
    """bayer2rgb ! coloreffects ! solarize driven through the element mirror on memories: the three elements only record
    themselves; the host read of the last memory launches ONE fused kernel. Bit-exact against the oracle chain
    (gstbayer2rgb.c:387-451, gstcoloreffects.c:303-359, gstsolarize.c:286-339), at a test size and at 7680x4320."""
    w, h = size
    src = frames.random_u8(rng, h, w)
    rgb = orc.bayer2rgb(src, w, h, "bggr", "BGRx")
    want = orc.solarize(orc.coloreffects(rgb, w, h, "BGRx", preset).view(np.uint32)).view(np.uint8).reshape(h, 4 * w)
    e1, e2, e3 = ctx.element("bayer2rgb"), ctx.element("coloreffects"), ctx.element("solarize")
    e1.set_caps("bggr", "BGRx", w, h); e2.set_caps("BGRx", "BGRx", w, h); e3.set_caps("BGRx", "BGRx", w, h)
    e2.set_property("preset", preset)
    m0, m1, m2 = ctx.memory(w * h), ctx.memory(4 * w * h), ctx.memory(4 * w * h)
    c0, l0 = ctx.transfer_counts(), ctx.launch_count()
    m0.write(src)
    e1.transform_mem(m0, m1)
    e2.transform_mem(m1, m1)                       # transform_frame_ip
    e3.transform_mem(m1, m2)
    assert ctx.launch_count() == l0 and counts_delta(ctx, c0) == (0, 0, 0, 0)         # nothing has run yet
    assert (m1.pending_stages, m2.pending_stages) == (2, 3)
    m0.close()                                     # the chain holds its own reference to the source
    got = m2.read().reshape(h, 4 * w)
    assert ctx.launch_count() - l0 == 1 and ctx.last_kernel().startswith("bayer2rgb_tma_fused"), ctx.last_kernel()
    assert counts_delta(ctx, c0) == (1, w * h, 1, 4 * w * h)
    assert np.array_equal(got, want), (preset, np.argwhere(got != want)[:4])
    # the intermediate memory can still be read: its own (shorter) chain runs then
    mid = m1.read().reshape(h, 4 * w)
    assert np.array_equal(mid, orc.coloreffects(rgb, w, h, "BGRx", preset))
    for m in (m1, m2):
        m.close()


def test_chain_without_deferral_gives_the_same_bytes_in_three_launches(ctx, vf, orc, rng, monkeypatch):
    w, h = 512, 96
    src = frames.random_u8(rng, h, w)
    rgb = orc.bayer2rgb(src, w, h, "bggr", "RGBx")
    want = orc.solarize(orc.coloreffects(rgb, w, h, "RGBx", "heat").view(np.uint32)).view(np.uint8).reshape(h, 4 * w)
    monkeypatch.setenv("B200VF_NO_DEFER", "1")
    e1, e2, e3 = ctx.element("bayer2rgb"), ctx.element("coloreffects"), ctx.element("solarize")
    e1.set_caps("bggr", "RGBx", w, h); e2.set_caps("RGBx", "RGBx", w, h); e3.set_caps("RGBx", "RGBx", w, h)
    e2.set_property("preset", "heat")
    m0, m1, m2 = ctx.memory(w * h), ctx.memory(4 * w * h), ctx.memory(4 * w * h)
    m0.write(src)
    c0, l0 = ctx.transfer_counts(), ctx.launch_count()
    e1.transform_mem(m0, m1); e2.transform_mem(m1, m1); e3.transform_mem(m1, m2)
    assert ctx.launch_count() - l0 == 3
    got = m2.read().reshape(h, 4 * w)
    assert counts_delta(ctx, c0) == (1, w * h, 1, 4 * w * h)                          # still one transfer each way
    assert np.array_equal(got, want)


def test_lut_elements_alone_compose_into_one_launch(ctx, vf, orc, rng):
    """burn ! dodge ! solarize on a frame that is already in HBM: three per-channel elements, one lut4 launch"""
    w, h = 640, 48
    fr = frames.random_u8(rng, h, 4 * w)
    want = orc.solarize(orc.dodge(orc.burn(fr.view(np.uint32), 175)), 127, 50, 185).view(np.uint8).reshape(h, 4 * w)
    els = [ctx.element(n) for n in ("burn", "dodge", "solarize")]
    for e in els:
        e.set_caps("BGRx", "BGRx", w, h)
    mems = [ctx.memory(4 * w * h) for _ in range(4)]
    mems[0].write(fr)
    l0 = ctx.launch_count()
    for i, e in enumerate(els):
        e.transform_mem(mems[i], mems[i + 1])
    assert ctx.launch_count() == l0 and mems[3].pending_stages == 3
    got = mems[3].read().reshape(h, 4 * w)
    assert ctx.launch_count() - l0 == 1 and ctx.last_kernel().startswith("lut4")
    assert np.array_equal(got, want)


def test_element_that_cannot_join_flushes_the_chain(ctx, vf, orc, rng):
    """bayer2rgb ! burn ! dilate: dilate is a stencil - the pending (bayer2rgb + burn) chain is launched when dilate maps
    its input, then dilate runs; a luma preset after a LUT cannot fold either (the fused kernel applies it first)"""
    w, h = 256, 64
    src = frames.random_u8(rng, h, w)
    rgb = orc.bayer2rgb(src, w, h, "rggb", "BGRx")
    burned = orc.burn(rgb.view(np.uint32), 100)
    want = orc.dilate(burned.reshape(h, w), False).view(np.uint8).reshape(h, 4 * w)
    e1, e2, e3 = ctx.element("bayer2rgb"), ctx.element("burn"), ctx.element("dilate")
    e1.set_caps("rggb", "BGRx", w, h); e2.set_caps("BGRx", "BGRx", w, h); e3.set_caps("BGRx", "BGRx", w, h)
    e2.set_property("adjustment", 100)
    m = [ctx.memory(w * h)] + [ctx.memory(4 * w * h) for _ in range(3)]
    m[0].write(src)
    l0 = ctx.launch_count()
    e1.transform_mem(m[0], m[1]); e2.transform_mem(m[1], m[2])
    assert ctx.launch_count() == l0
    e3.transform_mem(m[2], m[3])
    assert ctx.launch_count() - l0 == 2
    assert np.array_equal(m[3].read().reshape(h, 4 * w), want)
    # solarize, then a luma preset in place: the preset cannot join behind a LUT -> the chain is launched, then coloreffects
    e4, e5 = ctx.element("solarize"), ctx.element("coloreffects")
    e4.set_caps("BGRx", "BGRx", w, h); e5.set_caps("BGRx", "BGRx", w, h)
    e5.set_property("preset", "sepia")
    m2 = [ctx.memory(w * h)] + [ctx.memory(4 * w * h) for _ in range(2)]
    m2[0].write(src)
    e1.transform_mem(m2[0], m2[1]); e4.transform_mem(m2[1], m2[2]); e5.transform_mem(m2[2], m2[2])
    want2 = orc.coloreffects(orc.solarize(rgb.view(np.uint32)).view(np.uint8).reshape(h, 4 * w), w, h, "BGRx", "sepia")
    assert np.array_equal(m2[2].read().reshape(h, 4 * w), want2)


def test_out_of_place_elements_reject_aliased_memories(ctx, vf):
    e = ctx.element("dilate")
    e.set_caps("BGRx", "BGRx", 64, 16)
    m = ctx.memory(64 * 16 * 4)
    with pytest.raises(vf.B200vfError) as err:
        e.transform_mem(m, m)
    assert err.value.status == vf.E_INVAL
    g = ctx.element("gaussianblur")
    g.set_caps("AYUV", "AYUV", 64, 16)
    d = ctx.alloc(64 * 16 * 4)
    with pytest.raises(vf.B200vfError):
        g.transform_device(d, d)


def test_two_streams_are_ordered_through_the_memory(ctx, vf, orc, rng, monkeypatch):
    """an element writes a memory on stream A (behind a long blur, so that it is still queued), another element reads it
    on stream B straight away: the memory orders B after its last use on A (GstCudaMemory synchronises instead)"""
    import torch
    monkeypatch.setenv("B200VF_NO_DEFER", "1")                 # launch every element as it comes
    w, h = 3840, 2160
    fr = frames.random_u8(rng, h, 4 * w)
    want = orc.dilate(orc.burn(fr.view(np.uint32), 175).reshape(h, w), False).view(np.uint8).reshape(h, 4 * w)
    burn, dil = ctx.element("burn"), ctx.element("dilate")
    burn.set_caps("BGRx", "BGRx", w, h); dil.set_caps("BGRx", "BGRx", w, h)
    sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
    k, ks = vf.gauss_kernel(5.0)
    big = ctx.upload(frames.random_u8(rng, 4 * h, 4 * w))
    big_out = ctx.alloc(4 * h * 4 * w + 64)
    for rep in range(3):
        m0, m1, m2 = ctx.memory(4 * w * h), ctx.memory(4 * w * h), ctx.memory(4 * w * h)
        m0.write(fr)
        ctx.gaussblur(big, big_out, w, h, 4 * w, 1, k, ks, nframes=4, stream=sa.cuda_stream)   # ~0.5 ms of work ahead on A
        burn.transform_mem(m0, m1, stream=sa.cuda_stream)
        dil.transform_mem(m1, m2, stream=sb.cuda_stream)                                           # B: no wait of its own
        got = m2.read().reshape(h, 4 * w)
        assert np.array_equal(got, want), rep
        for m in (m0, m1, m2):
            m.close()
