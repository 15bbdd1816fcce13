"""videofiltersbad plugin: zebrastripe, videodiff (luma loop), scenechange (SAD + decision)
through the C-ABI, bit-exact against the oracle."""
import numpy as np
import pytest

import frames

pytestmark = pytest.mark.gpu

# widths: < 4 (bytewise tail only), not a multiple of 4, pitch not 16-byte aligned (642 -> 644), aligned, > one CTA
SIZES = [(3, 5), (7, 5), (22, 13), (130, 21), (642, 33), (1024, 40), (3840, 18)]


def up(ctx, a):
    return ctx.upload(np.ascontiguousarray(a))


@pytest.mark.parametrize("w,h", SIZES)
def test_zebrastripe_planar(ctx, orc, rng, w, h):
    st = frames.round_up_4(w)
    fr = frames.random_u8(rng, h, st)
    for thr in (0, 37, 90, 100):
        for t in (0, 1, 6):
            d = up(ctx, fr)
            ctx.zebrastripe(d, 1, st, w, h, threshold=thr, t=t)
            got = ctx.download(d, fr.size).reshape(h, st)
            want = orc.zebrastripe(fr, w, h, thr, t)
            assert np.array_equal(got, want), (w, h, thr, t, ctx.last_kernel(), np.argwhere(got != want)[:4])


@pytest.mark.parametrize("w,h", [(7, 5), (130, 21), (640, 24)])
def test_zebrastripe_packed_formats(ctx, orc, rng, w, h):
    """YUY2 (luma at even bytes), UYVY (odd bytes), AYUV (byte 1 of 4): only the luma bytes may change"""
    for ps, off, stride in [(2, 0, frames.round_up_4(2 * w)), (2, 1, frames.round_up_4(2 * w)), (4, 1, 4 * w)]:
        fr = frames.random_u8(rng, h, stride)
        d = up(ctx, fr)
        ctx.zebrastripe(d.ptr + off, ps, stride, w, h, threshold=45, t=3)
        got = ctx.download(d, fr.size).reshape(h, stride)
        want = orc.zebrastripe(fr, w, h, 45, 3, ps, off)
        assert np.array_equal(got, want), (w, h, ps, off)


def test_zebrastripe_batch_counts_t_per_frame(ctx, orc, rng):
    w, h, n = 256, 32, 5
    fr = rng.integers(0, 256, (n, h, w), dtype=np.uint8)
    d = up(ctx, fr)
    ctx.zebrastripe(d, 1, w, w, h, threshold=50, t=2, nframes=n)
    got = ctx.download(d, fr.size).reshape(n, h, w)
    for f in range(n):
        assert np.array_equal(got[f], orc.zebrastripe(fr[f], w, h, 50, 2 + f)), f


@pytest.mark.parametrize("w,h", SIZES)
def test_videodiff_luma(ctx, orc, rng, w, h):
    st = frames.round_up_4(w)
    old = frames.random_u8(rng, h, st)
    new = old.copy()
    m = rng.random((h, st)) < 0.5
    new[m] = (old[m].astype(int) + rng.integers(-30, 31, int(m.sum()))).clip(0, 255).astype(np.uint8)
    for thr, t in [(10, 0), (10, 5), (0, 1), (254, 0), (255, 0), (300, 2), (-1, 3)]:
        d_out = ctx.alloc(old.size)
        ctx.videodiff_luma(up(ctx, old), up(ctx, new), d_out, st, w, h, threshold=thr, t=t)
        got = ctx.download(d_out, old.size).reshape(h, st)[:, :w]
        want = orc.videodiff_luma(old, new, w, h, thr, t)[:, :w]
        assert np.array_equal(got, want), (w, h, thr, t, np.argwhere(got != want)[:4])


@pytest.mark.parametrize("w,h", SIZES)
def test_sad_u8(ctx, orc, rng, w, h):
    st = frames.round_up_4(w)
    n = 3
    a = rng.integers(0, 256, (n, h, st), dtype=np.uint8)
    b = rng.integers(0, 256, (n, h, st), dtype=np.uint8)
    b[1] = a[1]                                            # identical pair: 0
    d_sums = ctx.alloc(4 * n)
    ctx.sad_u8(up(ctx, a), up(ctx, b), st, w, h, d_sums, nframes=n)
    got = ctx.download(d_sums, 4 * n).view(np.uint32)
    want = [orc.sad_u8(a[i], b[i], w, h) for i in range(n)]
    assert list(got) == want, (w, h)


def test_sad_u8_wraps_like_the_32_bit_accumulator(ctx, orc):
    """8K luma planes of 0 and 255 differ by 255 * 33.2 M = 8.46e9 > 2^32: the reference's orc_uint32 wraps,
    and so must the device sum (size-independent property: the sum is linear in the number of identical rows)"""
    w, h = 7680, 4320
    a = np.zeros((h, w), np.uint8)
    b = np.full((h, w), 255, np.uint8)
    d_sums = ctx.alloc(4)
    ctx.sad_u8(up(ctx, a), up(ctx, b), w, w, h, d_sums)
    got = int(ctx.download(d_sums, 4).view(np.uint32)[0])
    assert got == (255 * w * h) % (1 << 32)
    assert got == orc.sad_u8(a[:1], b[:1], w, 1) * h % (1 << 32)


def test_scenechange_scores_from_device_sads(ctx, vf, orc, rng):
    """the element end to end on a synthetic clip: luma planes drift slowly, with two hard cuts; the score is
    SAD / (w * h) from the device, the decision the library's host state machine; both equal the oracle's"""
    w, h, n = 320, 180, 40
    base = rng.integers(0, 256, (h, w), dtype=np.uint8)
    clip = []
    for f in range(n):
        if f in (17, 31):
            base = rng.integers(0, 256, (h, w), dtype=np.uint8)
        noise = rng.integers(-3, 4, (h, w))
        base = (base.astype(int) + noise).clip(0, 255).astype(np.uint8)
        clip.append(base)
    clip = np.stack(clip)
    d = up(ctx, clip)
    d_sums = ctx.alloc(4 * (n - 1))
    # pair f = (frame f, frame f + 1): one batched launch, the old frame is the previous one of the same slab
    ctx.sad_u8(d, d.ptr + h * w, w, w, h, d_sums, nframes=n - 1, frame_stride=h * w)
    sads = ctx.download(d_sums, 4 * (n - 1)).view(np.uint32)
    assert list(sads) == [orc.sad_u8(clip[f], clip[f + 1], w, h) for f in range(n - 1)]
    scores = [float(s) / (w * h) for s in sads]
    st = vf.SceneChange()
    got = [st.update(x) for x in scores]
    assert got == orc.scenechange_run(scores)
    assert [i + 1 for i, c in enumerate(got) if c] == [17, 31]


# ---------------------------------------------------------------- through the element mirror
def _i420(rng, w, h, n):
    ru = lambda v, a: (v + a - 1) // a * a
    s0, h2 = ru(w, 4), ru(h, 2)
    size = s0 * h2 + 2 * ru(ru(w, 2) // 2, 4) * (h2 // 2)
    return rng.integers(0, 256, (n, size), dtype=np.uint8), s0


def test_zebrastripe_element_i420_and_ayuv(ctx, orc, rng):
    """element surface: `threshold` property, frame counter t across calls, chroma planes untouched"""
    w, h, n = 70, 34, 3
    fr, s0 = _i420(rng, w, h, n)
    e = ctx.element("zebrastripe")
    e.set_caps("I420", "I420", w, h)
    e.set_property("threshold", 40)
    for call in range(2):                                     # t keeps counting: frames 0..2, then 3..5
        got = e.transform(fr, n).reshape(n, -1)
        for f in range(n):
            want = fr[f].copy()
            want[: s0 * h] = orc.zebrastripe(fr[f][: s0 * h].reshape(h, s0), w, h, 40, call * n + f).reshape(-1)
            assert np.array_equal(got[f], want), (call, f)
    e.close()
    ay = rng.integers(0, 256, (1, h, 4 * w), dtype=np.uint8)
    e = ctx.element("zebrastripe")
    e.set_caps("AYUV", "AYUV", w, h)
    got = e.transform(ay, 1).reshape(h, 4 * w)
    assert np.array_equal(got, orc.zebrastripe(ay[0], w, h, 90, 0, 4, 1))
    e.close()


def test_videodiff_element_sequence(ctx, orc, rng):
    """first frame passes through, every later one is compared with its predecessor - across a device batch and
    across calls (host path, one frame per call)"""
    w, h, n = 66, 30, 5
    fr, s0 = _i420(rng, w, h, n)
    for f in range(1, n):                                     # luma drifts a little, with a few big changes
        y = fr[f - 1][: s0 * h].astype(int) + rng.integers(-12, 13, s0 * h)
        fr[f][: s0 * h] = y.clip(0, 255).astype(np.uint8)
    want = fr.copy()
    for f in range(1, n):
        want[f][: s0 * h] = orc.videodiff_luma(fr[f - 1][: s0 * h].reshape(h, s0), fr[f][: s0 * h].reshape(h, s0), w, h, 10, 0).reshape(-1)
    # rows of the luma plane beyond `w` are padding: the reference loop never writes them, the kernel works on the
    # pixels only, the rest of the plane is the copy of the input
    e = ctx.element("videodiff")
    e.set_caps("I420", "I420", w, h)
    d_in, d_out = ctx.upload(fr), ctx.alloc(fr.size)
    e.transform_device(d_in, d_out, n)                        # one batch
    got = ctx.download(d_out, fr.size).reshape(n, -1)
    luma = lambda a: a[:, : s0 * h].reshape(len(a), h, s0)[:, :, :w]
    assert np.array_equal(luma(got), luma(want)) and np.array_equal(got[:, s0 * h:], fr[:, s0 * h:])
    e.close()
    e = ctx.element("videodiff")
    e.set_caps("I420", "I420", w, h)
    got2 = np.stack([e.transform(fr[f], 1) for f in range(n)])    # frame by frame through the host path
    assert np.array_equal(luma(got2), luma(want)) and np.array_equal(got2[:, s0 * h:], fr[:, s0 * h:])
    e.close()


def test_scenechange_element_events(ctx, vf, orc, rng):
    w, h, n = 160, 90, 30
    clip, s0 = _i420(rng, w, h, n)
    base = clip[0][: s0 * h].copy()
    for f in range(n):
        if f in (12, 21):
            base = rng.integers(0, 256, s0 * h, dtype=np.uint8)
        base = (base.astype(int) + rng.integers(-3, 4, s0 * h)).clip(0, 255).astype(np.uint8)
        clip[f][: s0 * h] = base
    scores = [orc.sad_u8(clip[f][: s0 * h].reshape(h, s0), clip[f + 1][: s0 * h].reshape(h, s0), w, h) / (w * h) for f in range(n - 1)]
    want = [False] + orc.scenechange_run(scores)              # frame 0 only primes the element
    e = ctx.element("scenechange")
    e.set_caps("I420", "I420", w, h)
    d = ctx.upload(clip)
    e.transform_device(d, d, 10)                              # in place, three batches: state carries over
    ev = e.last_events()
    e.transform_device(d.ptr + 10 * clip.shape[1], d.ptr + 10 * clip.shape[1], 13)
    ev += e.last_events()
    e.transform_device(d.ptr + 23 * clip.shape[1], d.ptr + 23 * clip.shape[1], 7)
    ev += e.last_events()
    assert ev == want and [i for i, c in enumerate(ev) if c] == [12, 21]
    assert np.array_equal(ctx.download(d, clip.size).reshape(clip.shape), clip)      # passthrough
    e.close()
    e = ctx.element("scenechange")
    e.set_caps("I420", "I420", w, h)
    out = e.transform(clip, n)                                # host path, frame by frame
    assert e.last_events() == want and np.array_equal(out.reshape(clip.shape), clip)
    e.close()


# ---------------------------------------------------------------- smooth (gst/smooth)
def _plateaus(rng, h, st):
    return ((rng.integers(0, 256, (h, st)) // 32) * 32 + rng.integers(0, 12, (h, st))).astype(np.uint8)


@pytest.mark.parametrize("w,h", [(1, 1), (1, 5), (5, 1), (4, 2), (22, 13), (70, 40), (130, 37), (642, 50)])
def test_smooth_plane(ctx, orc, rng, w, h):
    """smooth_filter incl. its quirks: reference sample one row above the window centre, asymmetric rows, row 0
    written twice, last row never written (it keeps what the destination held: 77 here)"""
    st = frames.round_up_4(w)
    src = _plateaus(rng, h, st)
    for tol, fs in [(8, 3), (0, 3), (1, 1), (-5, 2), (300, 1), (8, 0), (8, -1), (20, 8), (-40000, 2)]:      # beyond |tolerance| ~ 46000 the reference's int product overflows (undefined)
        d_dst = up(ctx, np.full((h, st), 77, np.uint8))
        ctx.smooth_plane(up(ctx, src), d_dst, st, w, h, tolerance=tol, filtersize=fs)
        got = ctx.download(d_dst, h * st).reshape(h, st)
        want = orc.smooth_plane(src, w, h, tol, fs, 77)
        assert np.array_equal(got[:, :w], want[:, :w]), (w, h, tol, fs, np.argwhere(got[:, :w] != want[:, :w])[:4])
        assert (got[:, w:] == 77).all()                      # padding columns are not touched
    with pytest.raises(Exception):
        ctx.smooth_plane(up(ctx, src), up(ctx, src), st, w, h, filtersize=9)       # > supported window


def test_smooth_element_i420(ctx, orc, rng):
    w, h, n = 70, 34, 2
    s0, h2 = 72, 34
    cw, ch, s1 = 35, 17, 36
    size = s0 * h2 + 2 * s1 * ch
    fr = np.stack([_plateaus(rng, 1, size)[0] for _ in range(n)])
    e = ctx.element("smooth")
    e.set_caps("I420", "I420", w, h)
    for luma_only in (True, False):
        e.set_property("luma-only", 1 if luma_only else 0)
        d_in, d_out = up(ctx, fr), up(ctx, np.full_like(fr, 77))
        e.transform_device(d_in, d_out, n)
        got = ctx.download(d_out, fr.size).reshape(n, size)
        for f in range(n):
            y = orc.smooth_plane(fr[f][: s0 * h].reshape(h, s0), w, h, 8, 3, 77)
            assert np.array_equal(got[f][: s0 * h].reshape(h, s0)[:, :w], y[:, :w]), (luma_only, f)
            for off in (s0 * h2, s0 * h2 + s1 * ch):
                plane = fr[f][off: off + s1 * ch].reshape(ch, s1)
                g = got[f][off: off + s1 * ch].reshape(ch, s1)
                if luma_only:
                    assert np.array_equal(g, plane)          # gst_video_frame_copy_plane
                else:
                    assert np.array_equal(g[:, :cw], orc.smooth_plane(plane, cw, ch, 8, 3, 77)[:, :cw])
    e.set_property("active", 0)
    out = e.transform(fr, n).reshape(n, size)
    assert np.array_equal(out, fr)                            # inactive: gst_video_frame_copy
    e.close()


# ---------------------------------------------------------------- videoanalyse (gst/videosignal)
@pytest.mark.parametrize("w,h", SIZES)
def test_videoanalyse_moments(ctx, vf, orc, rng, w, h):
    st = frames.round_up_4(w)
    n = 3
    a = rng.integers(0, 256, (n, h, st), dtype=np.uint8)
    a[1] = 255
    d_sums = ctx.alloc(16 * n)
    ctx.luma_moments(up(ctx, a), st, w, h, d_sums, nframes=n)
    got = ctx.download(d_sums, 16 * n).view(np.uint64).reshape(n, 2)
    for f in range(n):
        x = a[f][:, :w].astype(np.uint64)
        assert (int(got[f, 0]), int(got[f, 1])) == (int(x.sum()), int((x * x).sum())), (w, h, f)
        assert vf.videoanalyse_finish(int(got[f, 0]), int(got[f, 1]), w, h) == orc.videoanalyse(a[f], w, h)


def test_videoanalyse_8k_needs_64_bit_sums(ctx, vf):
    w, h = 7680, 4320
    a = np.full((h, w), 255, np.uint8)
    d_sums = ctx.alloc(16)
    ctx.luma_moments(up(ctx, a), w, w, h, d_sums)
    got = ctx.download(d_sums, 16).view(np.uint64)
    assert (int(got[0]), int(got[1])) == (255 * w * h, 255 * 255 * w * h)      # 8.5e9 and 2.2e12
    assert vf.videoanalyse_finish(int(got[0]), int(got[1]), w, h) == (1.0, 0.0)


# ---------------------------------------------------------------- videosignal: simplevideomark / simplevideomarkdetect
def _mark_cases(rng, n):
    for _ in range(n):
        w, h = int(rng.integers(8, 200)), int(rng.integers(4, 90))
        ps = int(rng.choice([1, 2, 4]))
        st = frames.round_up_4(w * ps + int(rng.integers(0, 3)) * 4)
        kw = dict(pw=int(rng.integers(1, 12)), ph=int(rng.integers(1, h + 8)), pc=int(rng.integers(0, 7)), pdc=int(rng.integers(0, 12)),
                  left=int(rng.integers(0, w + 4)), bottom=int(rng.integers(0, h + 3)))
        yield w, h, ps, st, kw, int(rng.integers(0, 1 << 12))


def test_simplevideomark_draws_the_references_boxes(ctx, vf, orc, rng):
    """gst_video_mark_yuv (gstsimplevideomark.c:348-462): boxes clipped at the right / top edge, offsets beyond the
    frame (nothing drawn), zero counts, packed layouts (pixel stride 2 and 4) - 120 random configurations + defaults"""
    cases = list(_mark_cases(rng, 120)) + [(640, 480, 1, 640, dict(pw=4, ph=16, pc=4, pdc=5, left=0, bottom=0), 10)]
    for (w, h, ps, st, kw, data) in cases:
        fr = rng.integers(0, 256, (h, st), dtype=np.uint8)
        p = vf.VideoMarkParams(kw["pw"], kw["ph"], kw["pc"], kw["pdc"], kw["left"], kw["bottom"])
        d = up(ctx, fr)
        ctx.videomark_draw(d, ps, st, w, h, p, pattern_data=data)
        got = ctx.download(d, h * st).reshape(h, st)
        want = orc.videomark(fr, ps, w, h, data=data, **kw)
        assert np.array_equal(got, want), (w, h, ps, kw, np.argwhere(got != want)[:4])


def test_simplevideomarkdetect_reads_what_simplevideomark_wrote(ctx, vf, orc, rng):
    """box sums on the GPU + the reference's decisions on the host == gst_video_detect_yuv, on marked and unmarked frames,
    from both states of in_pattern (a pattern that disappears posts have-pattern = false once)"""
    for (w, h, ps, st, kw, data) in list(_mark_cases(rng, 80)) + [(640, 480, 1, 640, dict(pw=4, ph=16, pc=4, pdc=5, left=0, bottom=0), 21)]:
        # two rows of slack: where a clipped box is averaged over its FULL width the reference reads on into the next rows
        fr = np.zeros((h + 2, st), np.uint8)
        fr[:h] = rng.integers(0, 256, (h, st), dtype=np.uint8)
        marked = orc.videomark(fr, ps, w, h, data=data, **kw)
        p = vf.VideoMarkParams(kw["pw"], kw["ph"], kw["pc"], kw["pdc"], kw["left"], kw["bottom"])
        for frame in (fr, marked):
            sums = ctx.videomark_box_sums(up(ctx, frame[:h]), ps, st, w, h, p)[0]
            for ip in (False, True):
                got = vf.videomark_detect_decide(p, w, h, st, ps, sums, 0.5, 0.3, ip)
                want = orc.videomarkdetect(frame, ps, w, h, in_pattern=ip, **{k: v for k, v in kw.items()})
                # (the plane handed to the GPU ends at row h: samples the reference reads in the slack rows are zero on both sides)
                assert got == want, (w, h, ps, kw, ip, got, want)
    # the default mark on a default-sized frame carries its data
    fr = np.full((480, 640), 128, np.uint8)
    marked = orc.videomark(fr, 1, 640, 480, data=21)
    p = vf.VideoMarkParams()
    assert vf.videomark_detect_decide(p, 640, 480, 640, 1, ctx.videomark_box_sums(up(ctx, marked), 1, 640, 640, 480, p)[0]) == (True, True, 21)


def test_videosignal_elements(ctx, vf, orc, rng):
    """simplevideomark ! simplevideomarkdetect and videoanalyse through the element mirror (I420 and UYVY), on memories:
    the frame stays in HBM, the detector / analyser bring back a few numbers"""
    w, h = 320, 240
    for fmt, ps, off in (("I420", 1, 0), ("UYVY", 2, 1)):
        mark, det = ctx.element("simplevideomark"), ctx.element("simplevideomarkdetect")
        for e in (mark, det):
            e.set_caps(fmt, fmt, w, h)
            e.set_property("pattern-width", 6)
            e.set_property("left-offset", 10)
        mark.set_property("pattern-data", 19)
        n_in, _ = mark.unit_size()
        fr = rng.integers(0, 256, n_in, dtype=np.uint8)
        m = ctx.memory(n_in)
        m.write(fr)
        c0 = ctx.transfer_counts()
        mark.transform_mem(m, m)
        det.transform_mem(m, m)
        assert det.last_values() == [1.0, 1.0, 19.0]
        c1 = ctx.transfer_counts()
        assert (c1[0] - c0[0], c1[2] - c0[2]) == (1, 0)         # one upload, nothing downloaded through the memory
        got = m.read()
        st = frames.round_up_4(w * ps)
        plane = fr[: st * h].reshape(h, st).copy()
        want = orc.videomark(plane[:, off:] if off else plane, ps, w, h, pw=6, left=10, data=19) if not off else None
        if not off:
            assert np.array_equal(got[: st * h].reshape(h, st), want) and np.array_equal(got[st * h:], fr[st * h:])
        mark.set_property("enabled", False)
        m.write(fr)
        mark.transform_mem(m, m)
        det.transform_mem(m, m)
        assert det.last_values()[:2] == [1.0, 0.0]                  # the pattern disappeared: one message, have-pattern false
        det.transform_mem(m, m)
        assert det.last_values()[0] == 0.0                          # ... and none after that
    va = ctx.element("videoanalyse")
    va.set_caps("I420", "I420", w, h)
    fr = rng.integers(0, 256, va.unit_size()[0], dtype=np.uint8)
    out = va.transform(fr)
    assert np.array_equal(out, fr)
    avg, var = orc.videoanalyse(fr[: w * h].reshape(h, w), w, h)
    assert va.last_values() == [avg, var]
