"""CPU-only: the host side of the product (no kernels run): the C-ABI library loads and exports every
symbol include/b200vf.h declares, fails loudly without a device, and its host-side builders (LUTs,
gaussian taps, geometric maps, index resolution, shard partition, element surface) match the oracle."""
import json
import os
import re

import numpy as np
import pytest

import oracle
import refprops

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = np.load(os.path.join(ROOT, "tests", "golden", "golden_small.npz"))


def test_library_exports_every_declared_symbol(vf):
    hdr = open(os.path.join(ROOT, "include", "b200vf.h")).read()
    declared = set(re.findall(r"\b(b200vf_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"b200vf_status"}
    assert declared, "no declarations parsed"
    missing = [s for s in sorted(declared) if not hasattr(vf.lib, s)]
    assert not missing, missing
    assert not vf.MISSING
    assert set(vf.declared_symbols()) == declared        # the binding covers the whole header, nothing more


def test_no_device_fails_loudly_not_silently(vf):
    try:
        c = vf.Context(0)
    except vf.B200vfError as e:
        assert e.status == vf.E_NO_DEVICE                  # CPU box: no fallback path exists
        assert "no CPU path" in str(e) or "sm_" in str(e)
    else:
        c.close()                                          # a GPU box: fine


def test_product_never_touches_the_oracle():
    """a product path that routes through oracle/ would void every parity claim"""
    pkg = os.path.join(ROOT, "gst-plugins-bad_b200")
    for dp, _, files in os.walk(pkg):
        if os.sep + "build" in dp or os.sep + "lib" in dp:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".c", "Makefile")):
                txt = open(os.path.join(dp, f), errors="replace").read()
                assert "import oracle" not in txt and "oracle/" not in txt.replace("the oracle/", ""), os.path.join(dp, f)


def px_all():
    b = np.arange(256, dtype=np.uint32)
    return b | (b << 8) | (b << 16) | (b << 24)


def apply_lut(lut, px):
    out = np.zeros_like(px)
    for c in range(4):
        out |= lut[c][(px >> (8 * c)) & 0xff].astype(np.uint32) << (8 * c)
    return out


def test_lut_builders_match_reference_arithmetic(vf, orc):
    px = G["px"]
    for adj in range(0, 257):
        assert np.array_equal(apply_lut(vf.lut_burn(adj), px_all()), orc.burn(px_all(), adj)), adj
    assert np.array_equal(apply_lut(vf.lut_dodge(), px), orc.dodge(px))
    for a, b in [(200, 1), (0, 0), (256, 256), (37, 255), (1, 128)]:
        assert np.array_equal(apply_lut(vf.lut_chromium(a, b), px), orc.chromium(px, a, b))
    for t in range(0, 257, 16):
        for s in range(0, 257, 32):
            for e in range(0, 257, 32):
                assert np.array_equal(apply_lut(vf.lut_solarize(t, s, e), px_all()), orc.solarize(px_all(), t, s, e)), (t, s, e)
    with pytest.raises(vf.B200vfError):
        vf.lut_burn(257)


def test_lut_compose(vf):
    a, b = vf.lut_burn(100), vf.lut_solarize()
    c = vf.lut_compose(a, b)
    px = G["px"]
    assert np.array_equal(apply_lut(c, px), apply_lut(b, apply_lut(a, px)))


def test_gauss_kernel_matches_reference(vf, orc):
    for s in np.linspace(-20, 20, 81):
        k, ks = vf.gauss_kernel(float(s))
        rk, rks = orc.gauss_kernel(float(s))
        assert np.array_equal(k.view(np.uint32), rk.view(np.uint32)), s
        assert np.array_equal(ks.view(np.uint32), rks.view(np.uint32)), s
    with pytest.raises(vf.B200vfError):
        vf.gauss_kernel(20.5)


def test_coloreffects_tables_are_the_reference_tables(vf):
    gold = oracle.coloreffects_tables()
    for name, idx in oracle.PRESETS.items():
        if name != "none":
            t, ml = vf.coloreffects_table(idx)
            assert np.array_equal(t, gold[name]) and ml == oracle.PRESET_MAP_LUMA[name]


def test_geometric_maps_match_golden(vf):
    w, h = 16, 12
    for el, plist in refprops.CASES.items():
        for i, props in enumerate(plist):
            m = vf.gt_build_map(el, w, h, props)
            g = G["gt_map_%s_%d" % (el, i)]
            same = (m.view(np.uint64) == g.view(np.uint64)) | (np.isnan(m) & np.isnan(g))
            assert same.all(), (el, props)


@pytest.mark.skipif(not oracle.have_ref(), reason="needs oracle/_ref")
def test_geometric_maps_match_reference_larger(vf):
    R = oracle.get("reference")
    for el, plist in refprops.CASES.items():
        for props in plist:
            for (w, h) in [(64, 48), (100, 75)]:
                m = vf.gt_build_map(el, w, h, props)
                r = R.gt_map(el, w, h, refprops.full(el, props))
                assert ((m.view(np.uint64) == r.view(np.uint64)) | (np.isnan(m) & np.isnan(r))).all(), (el, props, w, h)


def test_resolve_map_is_do_map(vf, orc):
    """index table + a numpy gather == the oracle's per-pixel do_map loop, all three off-edge policies"""
    w, h = 16, 12
    src = G["gt_src"]
    for el in refprops.CASES:
        m = G["gt_map_%s_0" % el]
        for oe, code in oracle.OFF_EDGE.items():
            idx = vf.gt_resolve_map(m, w, h, code)
            px = src.reshape(h * w, 4)
            out = np.where((idx.reshape(-1) >= 0)[:, None], px[np.maximum(idx.reshape(-1), 0)], 0).astype(np.uint8)
            assert np.array_equal(out.reshape(h, 4 * w), G["gt_out_%s_%s" % (el, oe)]), (el, oe)
    # the tunnel centre pixel divides 0/0: it must stay unmapped (-1)
    m = vf.gt_build_map("tunnel", 64, 48)
    assert np.isnan(m[24, 32]).all() and vf.gt_resolve_map(m, 64, 48, 1)[24, 32] == -1


def test_unknown_element_and_property(vf):
    with pytest.raises(vf.B200vfError) as e:
        vf.gt_build_map("diffuse", 8, 8)
    assert e.value.status == vf.E_UNSUPPORTED
    with pytest.raises(vf.B200vfError) as e:
        vf.gt_build_map("bulge", 8, 8, {"nope": 1})
    assert e.value.status == vf.E_PROPERTY


def test_shard_rows_partition(vf):
    for h in [480, 2160, 4320, 2161, 1080]:
        for n in [1, 2, 4, 8]:
            cover = 0
            for r in range(n):
                r0, rows = vf.shard_rows(h, r, n)
                assert r0 == cover and r0 % 2 == 0 and rows >= 4
                cover += rows
            assert cover == h
    with pytest.raises(vf.B200vfError):
        vf.shard_rows(16, 0, 8)


def _num(s):
    if s is None:
        return None
    s = str(s).split(" ")[0]
    return float(s)


def test_element_surface_matches_reference_api_dump(vf):
    """Factory names, property names, ranges and defaults == docs/plugins/gst_plugins_cache.json of the
    reference (fixture extracted by tests/golden/make_golden.py); pad-template formats too."""
    surf = json.load(open(os.path.join(ROOT, "tests", "golden", "element_surface.json")))
    for name, el in surf.items():
        e = vf.Element(None, name)                          # no device needed to inspect an element
        for pn, pd in el["properties"].items():
            if pd["type"] == "GValueArray":                 # perspective's 3x3 matrix: exposed as matrix-0..8
                assert [e.get_property("matrix-%d" % i) for i in range(9)] == [1, 0, 0, 0, 1, 0, 0, 0, 1]
                continue
            got = e.get_property(pn)
            d = pd.get("default")
            if d in ("true", "false"):
                assert got == (1.0 if d == "true" else 0.0), (name, pn)
            elif "(" in str(d):                             # enum: "none (0)"
                nick, val = str(d).split(" (")
                assert got == float(val.rstrip(")")), (name, pn)
                e.set_property(pn, nick)                    # the nick round-trips
                assert e.get_property(pn) == got
            else:
                assert got == pytest.approx(_num(d), rel=2e-6), (name, pn, got, d)   # the dump prints 6 significant digits
        for pn, pd in el["properties"].items():            # ranges in a second pass (marble's turbulence aliases y-scale)
            if pd["type"] == "GValueArray":
                continue
            lo, hi = _num(pd.get("min")), _num(pd.get("max"))
            if lo is not None and lo > -1e300:
                e.set_property(pn, lo)
                if pd["type"] in ("guint", "gint") or lo > 0:
                    with pytest.raises(vf.B200vfError):
                        e.set_property(pn, lo - 1)
            if hi is not None and hi < 1e300 and hi < 2147483647:
                e.set_property(pn, hi)
                with pytest.raises(vf.B200vfError):
                    e.set_property(pn, hi + 1)
        # pad-template formats: every format the reference lists negotiates, an unlisted one does not
        caps = el["pad-templates"]["sink"]
        fmts = re.findall(r"format: \{ ([^}]*) \}", caps) or re.findall(r"format: (\w+)", caps)
        fmts = [f.strip() for f in fmts[0].split(",")] if fmts and "," in fmts[0] else fmts
        if name == "bayer2rgb":
            src_fmts = [f.strip() for f in re.findall(r"format: \{ ([^}]*) \}", el["pad-templates"]["src"])[0].split(",")]
            for f in fmts:
                for o in src_fmts:
                    e.set_caps(f, o, 64, 48)
                    assert e.unit_size() == (64 * 48, 64 * 48 * 4)
            with pytest.raises(vf.B200vfError):
                e.set_caps("bggr", "RGB", 64, 48)
        elif name == "rgb2bayer":
            e.set_caps("ARGB", "rggb", 63, 48)
            assert e.unit_size() == (63 * 48 * 4, 64 * 48)
        else:
            for f in fmts:
                e.set_caps(f, f, 33, 17)
            with pytest.raises(vf.B200vfError):
                e.set_caps("v308", "v308", 32, 16)             # in no pad template of these plugins
        e.close()


def test_element_without_device_cannot_transform(vf):
    e = vf.Element(None, "burn")
    e.set_caps("BGRx", "BGRx", 8, 8)
    with pytest.raises(vf.B200vfError) as err:
        e.transform(np.zeros(8 * 8 * 4, np.uint8))
    assert err.value.status == vf.E_NO_DEVICE


def test_factory_introspection_matches_reference_api_dump(vf):
    """what the GLib shells register (gst/gstb200vf.c reads exactly this table): plugin, GType name and parent,
    klass, long name, description, author, controllable flags == the reference's docs/plugins/gst_plugins_cache.json"""
    surf = json.load(open(os.path.join(ROOT, "tests", "golden", "element_surface.json")))
    fac = vf.factories()
    assert set(fac) == set(surf)
    for name, f in fac.items():
        e = surf[name]
        assert f["plugin"] == e["plugin"] and f["type_name"] == e["hierarchy"][0] and f["parent_type_name"] == e["hierarchy"][1]
        assert f["klass"] == e["klass"] and f["long_name"] == e["long-name"] and f["description"] == e["description"]
        assert f["author"] == e["author"] and f["plugin_license"] == e["plugin-license"]
        assert f["in_place"] == (1 if name in ("coloreffects", "chromahold", "zebrastripe", "scenechange", "videoanalyse", "simplevideomark",
                                               "simplevideomarkdetect") else 0)   # transform_frame_ip
        props = {p["name"]: p for p in f["properties"]}
        for pn, pd in e["properties"].items():
            if pd["type"] == "GValueArray":
                assert all(not props["matrix-%d" % i]["controllable"] for i in range(9))
                continue
            assert props[pn]["controllable"] == bool(pd.get("controllable")), (name, pn)
            kind = {"guint": 0, "gint": 1, "gboolean": 2, "gdouble": 3, "guint64": 5}.get(pd["type"], 4)
            assert props[pn]["type"] == kind, (name, pn, pd["type"])
        caps = e["pad-templates"]["src" if name == "rgb2bayer" else "sink"] if name not in ("bayer2rgb",) else e["pad-templates"]["src"]
        if name == "rgb2bayer":
            caps = e["pad-templates"]["sink"]
        fm = re.findall(r"format: \{ ([^}]*) \}", caps)
        want = [x.strip() for x in fm[0].split(",")] if fm else re.findall(r"format: (\w+)", caps)
        assert sorted(f["formats"]) == sorted(want), (name, f["formats"], want)


def test_packed_index_round_trips(vf):
    """b200vf_gt_pack_index is lossless for every map / off-edge policy / ragged width, and codes smooth maps almost fully"""
    rng = np.random.default_rng(7)
    r64 = lambda n: (n + 63) // 64 * 64
    for (w, h) in [(300, 40), (128, 16), (100, 9), (257, 5), (5, 3)]:
        for el, plist in refprops.CASES.items():
            if el == "diffuse":
                continue
            m = vf.gt_build_map(el, w, h, plist[0])
            for off_edge in (0, 1, 2):
                idx = vf.gt_resolve_map(m, w, h, off_edge)
                packed, raw = vf.gt_pack_index(idx, w, h)
                assert np.array_equal(vf.gt_unpack_index(packed, w, h), idx), (el, w, h, off_edge)
                chunks = ((w + 7) // 8) * h
                assert 0 <= raw <= chunks
                assert packed.size == 64 + r64(chunks * 4) + r64(chunks * 8) + raw * 32
    # a smooth map is coded completely; its table shrinks ~2.6x
    w, h = 512, 64
    idx = vf.gt_resolve_map(vf.gt_build_map("fisheye", w, h), w, h, 1)
    packed, raw = vf.gt_pack_index(idx, w, h)
    assert raw == 0 and packed.size < idx.nbytes / 2.6
    # arbitrary tables (random targets, ignored pixels) survive too: everything goes raw
    idx = rng.integers(-1, w * h, (h, w), dtype=np.int32)
    packed, raw = vf.gt_pack_index(idx, w, h)
    assert raw == (w // 8) * h and np.array_equal(vf.gt_unpack_index(packed, w, h), idx)
    # steps at the coding limits (-8 .. +7 in both directions) are coded, one beyond is not
    base = 40 * w + 200
    row = base + np.cumsum(np.r_[0, np.tile([7, -8, 7 * w + 7, -8 * w - 8], 32)[:127]])
    idx = np.tile(np.r_[row, row, row, row].astype(np.int32), (h, 1))
    packed, raw = vf.gt_pack_index(idx, w, h)
    assert raw == 0 and np.array_equal(vf.gt_unpack_index(packed, w, h), idx)   # the jump back at each 128th pixel starts a chunk
    idx[3, 5] += 8 + 1                                   # dx = 7 -> 16 (and the next step -17)
    packed2, raw2 = vf.gt_pack_index(idx, w, h)
    assert raw2 == raw + 1 and np.array_equal(vf.gt_unpack_index(packed2, w, h), idx)
    with pytest.raises(vf.B200vfError):
        vf.gt_unpack_index(packed2[:-8], w, h)           # truncated blob


def test_yuv_frame_geometry_of_the_videofilters_elements(vf):
    """default GstVideoInfo sizes (gst-plugins-base video-info.c fill_planes; external to the reference tree, so
    written down here independently): pitches rounded up to 4, chroma planes per format"""
    ru = lambda n, a: (n + a - 1) // a * a
    def size(fmt, w, h):
        s0, h2 = ru(w, 4), ru(h, 2)
        if fmt in ("I420", "YV12"):
            return s0 * h2 + 2 * ru(ru(w, 2) // 2, 4) * (h2 // 2)
        if fmt == "Y444":
            return 3 * s0 * h
        if fmt == "Y42B":
            return (s0 + ru(w, 8)) * h
        if fmt == "Y41B":
            return (s0 + ru(w, 16) // 2) * h
        if fmt in ("NV12", "NV21"):
            return s0 * h2 + s0 * (h2 // 2)
        if fmt in ("YUY2", "UYVY"):
            return ru(2 * w, 4) * h
        return 4 * w * h
    # known values (e.g. gst_video_info_set_format (I420, 320, 240) -> 115200; odd sizes round as above)
    assert size("I420", 320, 240) == 115200 and size("NV12", 320, 240) == 115200 and size("Y42B", 320, 240) == 153600
    assert size("Y41B", 320, 240) == 115200 and size("Y444", 320, 240) == 230400 and size("YUY2", 320, 240) == 153600
    e = vf.Element(None, "zebrastripe")
    for fmt in ["I420", "Y444", "Y42B", "Y41B", "YUY2", "UYVY", "AYUV", "NV12", "NV21", "YV12"]:
        for (w, h) in [(320, 240), (33, 17), (7, 5), (1, 1)]:
            e.set_caps(fmt, fmt, w, h)
            assert e.unit_size() == (size(fmt, w, h), size(fmt, w, h)), (fmt, w, h)
    assert e.get_property("threshold") == 90
    e.close()
    for name in ("videodiff", "scenechange"):
        e = vf.Element(None, name)
        with pytest.raises(vf.B200vfError):
            e.set_caps("NV12", "NV12", 32, 16)                # planar Y formats only (gstvideodiff.c:48-52)
        e.close()


def test_smooth_division_trick_is_exact():
    """smooth_dp4a (csrc/videofilters.cu) divides sum / count as floor((sum + 0.5) * RN(1 / count)) in fp32; numpy's
    float32 arithmetic is the same IEEE arithmetic (RN reciprocal, RN multiply), so every (count, sum) the kernel can
    see - count <= (2*6+1)*(2*6+3)+1 = 196 for filter-size <= 6 (324 checked), sum <= 255 * count - is checked against
    integer division here."""
    np.seterr(over="ignore")
    for num in range(1, 325):
        total = np.arange(0, 255 * num + 1, dtype=np.int64)
        rcp = np.float32(1.0) / np.float32(num)
        # the kernel uses rcp.approx (within 2 ulp of 1 / num): every reciprocal within +-3 ulp must give the same quotient
        for ulps in range(-3, 4):
            r = (rcp.view(np.uint32) + np.uint32(ulps & 0xffffffff)).view(np.float32) if ulps else rcp
            got = ((total.astype(np.float32) + np.float32(0.5)) * r).astype(np.int64)
            assert np.array_equal(got, total // num), (num, ulps)


def test_fp64_adder_truncation_trick_is_cvttsd2si():
    """csrc/gt_resolve.cuh converts double -> int without F2I: t = x + 1.5 * 2^52 rounds x to the nearest integer (ties to
    even) and leaves it in the low word; one compare against t - 1.5 * 2^52 turns nearest into toward-zero. Restated in
    numpy (IEEE doubles, round-to-nearest-even like the GPU's DADD) and compared with C's (int) on ties, neighbours of
    integers, negatives, the int32 borders, NaN and infinities."""
    M = np.float64(6755399441055744.0)

    def d2i(x):
        x = np.asarray(x, np.float64)
        with np.errstate(invalid="ignore"):
            inrange = np.abs(x) < 2147483648.0
            xs = np.where(inrange, x, 0.0)
            t = xs + M
            r = (t.view(np.uint64) & np.uint64(0xffffffff)).astype(np.uint32)
            rn = t - M
            r = np.where((xs >= 0) & (rn > xs), r - np.uint32(1), r)        # uint32 arithmetic wraps, as on the device
            r = np.where((xs < 0) & (rn < xs), r + np.uint32(1), r)
        return np.where(inrange, r.astype(np.uint32).view(np.int32).astype(np.int64), -2147483648)

    rng = np.random.default_rng(3)
    ints = rng.integers(-2**31 + 2, 2**31 - 2, 4000).astype(np.float64)
    cases = [ints, ints + 0.5, ints - 0.5, np.nextafter(ints, np.inf), np.nextafter(ints, -np.inf),
             rng.uniform(-1e4, 1e4, 20000), rng.uniform(-2.2e9, 2.2e9, 20000),
             np.array([0.0, -0.0, 0.5, -0.5, 1.5, -1.5, 2.5, -2.5, 0.9999999999999999, -0.9999999999999999, 2147483647.0,
                       2147483647.5, 2147483648.0, -2147483648.0, -2147483648.5, -2147483649.0, 1e300, -1e300, np.inf, -np.inf,
                       np.nan, 5e-324, -5e-324])]
    x = np.concatenate(cases)
    with np.errstate(invalid="ignore"):
        ok = np.abs(x) < 2147483649.0
        want = np.where(ok & (np.trunc(np.where(ok, x, 0)) >= -2147483648.0) & (np.trunc(np.where(ok, x, 0)) <= 2147483647.0),
                        np.trunc(np.where(ok, x, 0)), -2147483648.0).astype(np.int64)
    assert np.array_equal(d2i(x), want), x[d2i(x) != want][:10]


def test_host_builders_dense_sweeps(vf, orc):
    """the host-side builders against the reference's own C over dense parameter sweeps: chromium's cos-table LUT for
    every fourth (edge-a, edge-b) pair plus the borders, solarize for 4000 random (threshold, start, end) triples
    incl. start > end, gaussian taps and prefix sums for 1500 random sigmas (bit-equal floats)"""
    rng = np.random.default_rng(11)
    allb = px_all()
    edges = sorted(set(range(0, 257, 4)) | {1, 255, 256})
    for a in edges:
        for b in (0, 1, 2, 127, 128, 255, 256, int(rng.integers(0, 257))):
            assert np.array_equal(apply_lut(vf.lut_chromium(a, b), allb), orc.chromium(allb, a, b)), (a, b)
    for t, s, e in rng.integers(0, 257, (4000, 3)):
        assert np.array_equal(apply_lut(vf.lut_solarize(int(t), int(s), int(e)), allb), orc.solarize(allb, int(t), int(s), int(e))), (t, s, e)
    for s in np.concatenate([rng.uniform(-20, 20, 1200), rng.uniform(-1.5, 1.5, 300)]):
        s = float(np.float32(s))
        k, ks = vf.gauss_kernel(s)
        rk, rks = orc.gauss_kernel(s)
        assert np.array_equal(k.view(np.uint32), rk.view(np.uint32)) and np.array_equal(ks.view(np.uint32), rks.view(np.uint32)), s


@pytest.mark.skipif(not oracle.have_ref(), reason="needs oracle/_ref")
def test_geometric_maps_match_reference_over_random_properties(vf):
    """every deterministic map function against the reference's own *_map (compiled from /root/reference) for seeded
    random properties over their ranges, at a size whose centre is not a pixel: the double maps must be bit-equal"""
    R = oracle.get("reference")
    rng = np.random.default_rng(5)
    w, h = 93, 61
    circle = {"x_center": (0, 1), "y_center": (0, 1), "radius": (0, 1)}
    ranges = {
        "fisheye": {}, "tunnel": dict(circle), "mirror": {"mode": (0, 4)},
        "bulge": dict(circle, zoom=(1, 20)), "stretch": dict(circle, intensity=(0, 1)),
        "square": {"width": (0, 1), "height": (0, 1), "zoom": (1, 20)},
        "circle": dict(circle, angle=(-3.2, 3.2), spread_angle=(0.2, 6.3), height=(1, 80)),
        "kaleidoscope": dict(circle, angle=(-3.2, 3.2), angle2=(-3.2, 3.2), sides=(2, 9)),
        "pinch": dict(circle, intensity=(-1, 1)), "rotate": {"angle": (-6.3, 6.3)},
        "sphere": dict(circle, refraction=(0.3, 3.0)), "twirl": dict(circle, angle=(-6.3, 6.3)),
        "waterripple": dict(circle, amplitude=(-40, 40), phase=(-7, 7), wavelength=(2, 60)),
        "perspective": {"matrix_%d" % i: (-2, 2) for i in range(9)},
    }
    for el, rr in ranges.items():
        for trial in range(5):
            props = {k: (float(np.floor(rng.uniform(*v))) if k in ("sides", "height", "mode") and el in ("kaleidoscope", "circle", "mirror")
                         else float(rng.uniform(*v))) for k, v in rr.items()}
            m = vf.gt_build_map(el, w, h, props)
            r = R.gt_map(el, w, h, refprops.full(el, props))
            assert ((m.view(np.uint64) == r.view(np.uint64)) | (np.isnan(m) & np.isnan(r))).all(), (el, props)
