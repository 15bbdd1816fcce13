#!/usr/bin/env python3
"""Generates the golden fixtures of tests/golden/ FROM THE REFERENCE ITSELF.

Run where /root/reference is mounted (this container):   python tests/golden/make_golden.py
  * golden_small.npz         - seeded inputs + outputs of the reference's own C (oracle/_ref, built by
                               oracle/build_ref.py from /root/reference) for every op of the hot path;
  * coloreffects_tables.npz  - the five 256x3 preset tables (data of gstcoloreffects.c:117-286);
  * golden_videofilters.npz  - the same for zebrastripe / videodiff / scenechange;
  * element_surface.json     - factory name -> properties (type/min/max/default) and pad-template formats,
                               extracted from the reference's docs/plugins/gst_plugins_cache.json.
The reference's tests hold no vectors for these elements, so these fixtures are what pins
the oracle port when /root/reference is absent (the GPU box)."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle          # noqa: E402
import refprops        # noqa: E402

REF = os.environ.get("B200VF_REFERENCE", "/root/reference")


def make_videofilters():
    """golden_videofilters.npz: the videofiltersbad plugin's loops through the reference's C"""
    R = oracle.get("reference")
    rng = np.random.default_rng(20261017)
    g = {}
    w, h = 22, 13
    st = oracle.round_up_4(w)
    a = rng.integers(0, 256, (h, st), dtype=np.uint8)
    b = a.copy()
    m = rng.random((h, st)) < 0.4
    b[m] = rng.integers(0, 256, int(m.sum()), dtype=np.uint8)
    g["luma_a"], g["luma_b"] = a, b
    for thr in (0, 50, 90, 100):
        for t in (0, 3):
            g["zebra_%d_%d" % (thr, t)] = R.zebrastripe(a, w, h, thr, t)
    yuy2 = rng.integers(0, 256, (h, oracle.round_up_4(2 * w)), dtype=np.uint8)
    ayuv = rng.integers(0, 256, (h, 4 * w), dtype=np.uint8)
    g["yuy2"], g["ayuv"] = yuy2, ayuv
    g["zebra_yuy2"] = R.zebrastripe(yuy2, w, h, 40, 2, 2, 0)
    g["zebra_uyvy"] = R.zebrastripe(yuy2, w, h, 40, 2, 2, 1)
    g["zebra_ayuv"] = R.zebrastripe(ayuv, w, h, 40, 2, 4, 1)
    for t in (0, 5):
        g["videodiff_%d" % t] = R.videodiff_luma(a, b, w, h, 10, t)
    g["sad"] = np.array([R.sad_u8(a, b, w, h)], np.uint32)
    scores = np.concatenate([rng.uniform(0, 8, 30), [70.0], rng.uniform(0, 8, 8), [33.0, 2, 2, 2, 2, 2, 90, 100],
                             rng.uniform(20, 60, 20)])
    # smooth (its own plugin): plateaus + noise so that the tolerance matters; rows never written keep the prefill 77
    sm = ((rng.integers(0, 256, (h, st)) // 32) * 32 + rng.integers(0, 12, (h, st))).astype(np.uint8)
    g["smooth_src"] = sm
    for tol, fs in [(8, 3), (0, 3), (-5, 2), (300, 1), (8, 0), (8, -1), (20, 8)]:
        g["smooth_%d_%d" % (tol, fs)] = R.smooth_plane(sm, w, h, tol, fs, 77)
    g["va"] = np.array([R.videoanalyse(a, w, h), R.videoanalyse(sm, w, h), R.videoanalyse(np.full((h, st), 255, np.uint8), w, h)], np.float64)
    g["sc_scores"] = scores
    g["sc_changes"] = np.array(R.scenechange_run(scores), np.uint8)
    np.savez_compressed(os.path.join(HERE, "golden_videofilters.npz"), **g)
    print("golden_videofilters.npz: %d arrays" % len(g))


def make_surface():
    """element_surface.json from the reference's own API dump"""
    cache = json.load(open(os.path.join(REF, "docs", "plugins", "gst_plugins_cache.json")))
    want = {"bayer": ["bayer2rgb", "rgb2bayer"], "gaudieffects": None, "coloreffects": None, "geometrictransform": None,
            "videofiltersbad": ["zebrastripe", "videodiff", "scenechange"], "smooth": None, "videosignal": None}
    surf = {}
    for plugin, only in want.items():
        for name, el in cache[plugin]["elements"].items():
            if only and name not in only:
                continue
            props = {}
            for pn, pd in el.get("properties", {}).items():
                if pn in ("name", "parent", "qos"):
                    continue
                props[pn] = {k: pd.get(k) for k in ("type", "min", "max", "default", "controllable") if k in pd}
            pads = {pn: pd.get("caps") for pn, pd in el.get("pad-templates", {}).items()}
            surf[name] = {"plugin": plugin, "klass": el.get("klass"), "hierarchy": el.get("hierarchy"), "properties": props,
                          "pad-templates": pads, "long-name": el.get("long-name"), "description": el.get("description"),
                          "author": el.get("author"), "rank": el.get("rank"),
                          "plugin-description": cache[plugin].get("description"), "plugin-license": cache[plugin].get("license")}
    json.dump(surf, open(os.path.join(HERE, "element_surface.json"), "w"), indent=1, sort_keys=True)
    return surf


def main():
    make_videofilters()
    R = oracle.get("reference")
    rng = np.random.default_rng(20260925)
    g = {}
    np.savez_compressed(os.path.join(HERE, "coloreffects_tables.npz"),
                        **{p: R.coloreffects_table(p)[0] for p in ["heat", "sepia", "xray", "xpro", "yellowblue"]})
    # bayer2rgb / rgb2bayer
    for (w, h) in [(4, 3), (6, 5), (16, 10)]:
        src = rng.integers(0, 256, (h, oracle.round_up_4(w)), dtype=np.uint8)
        g["bayer_src_%dx%d" % (w, h)] = src
        for f in oracle.BAYER_FORMATS:
            for o in ["RGBA", "BGRA", "ARGB", "ABGR"]:
                g["bayer_%dx%d_%s_%s" % (w, h, f, o)] = R.bayer2rgb(src, w, h, f, o)
    argb = rng.integers(0, 256, (7, 4 * 9), dtype=np.uint8)
    g["rgb2bayer_src"] = argb
    for f in oracle.BAYER_FORMATS:
        g["rgb2bayer_" + f] = R.rgb2bayer(argb, 9, 7, f)
    # gaussianblur
    fr = rng.integers(0, 256, (9, 4 * 12), dtype=np.uint8)
    g["gauss_src"] = fr
    for s in [-1.2, 0.3, 1.2, 5.0]:
        k, ks = R.gauss_kernel(s)
        g["gauss_kernel_%g" % s] = k
        g["gauss_ksum_%g" % s] = ks
        for p0 in (0, 1, 2):
            g["gauss_%g_p%d" % (s, p0)] = R.gaussblur(fr, 12, 9, s, p0)
    # point ops: every byte value in every channel + random pixels
    b = np.arange(256, dtype=np.uint32)
    px = np.concatenate([b | (b << 8) | (b << 16) | (b << 24), rng.integers(0, 2 ** 32, 256, dtype=np.uint32)])
    g["px"] = px
    for adj in [0, 1, 175, 256]:
        g["burn_%d" % adj] = R.burn(px, adj)
    g["dodge"] = R.dodge(px)
    for a, bb in [(200, 1), (0, 0), (256, 256), (37, 255)]:
        g["chromium_%d_%d" % (a, bb)] = R.chromium(px, a, bb)
    for f in [1, 2, 100, 175]:
        g["exclusion_%d" % f] = R.exclusion(px, f)
    for t, s, e in [(127, 50, 185), (50, 50, 185), (100, 100, 100), (127, 185, 50), (10, 200, 30)]:
        g["solarize_%d_%d_%d" % (t, s, e)] = R.solarize(px, t, s, e)
    dsrc = rng.integers(0, 2 ** 32, (8, 16), dtype=np.uint32)
    g["dilate_src"] = dsrc
    g["dilate_0"] = R.dilate(dsrc, False)
    g["dilate_1"] = R.dilate(dsrc, True)
    # coloreffects / chromahold
    for fmt in ["RGB", "BGRx", "ARGB", "AYUV"]:
        ps = 3 if fmt == "RGB" else 4
        w, h = 9, 5
        cf = rng.integers(0, 256, (h, oracle.round_up_4(w * ps)), dtype=np.uint8)
        g["ce_src_" + fmt] = cf
        for pr in oracle.PRESETS:
            g["ce_%s_%s" % (fmt, pr)] = R.coloreffects(cf, w, h, fmt, pr)
    cf = rng.integers(0, 256, (6, 4 * 10), dtype=np.uint8)
    g["ch_src"] = cf
    for i, (tgt, tol) in enumerate([((255, 0, 0), 30), ((128, 128, 128), 0), ((0, 200, 30), 180)]):
        g["ch_%d" % i] = R.chromahold(cf, 10, 6, "xRGB", tgt, tol)
    # geometrictransform: maps (as raw doubles) and one gather per off-edge policy
    w, h = 16, 12
    gsrc = rng.integers(0, 256, (h, 4 * w), dtype=np.uint8)
    g["gt_src"] = gsrc
    for el, plist in refprops.CASES.items():
        for i, props in enumerate(plist):
            m = R.gt_map(el, w, h, refprops.full(el, props))
            g["gt_map_%s_%d" % (el, i)] = m
            if i == 0:
                for oe in oracle.OFF_EDGE:
                    g["gt_out_%s_%s" % (el, oe)] = R.remap(gsrc, m, w, h, 4, oe, False)
    g["gt_out_fisheye_ayuv"] = R.remap(gsrc, g["gt_map_fisheye_0"] * 1.7 - 9.0, w, h, 4, "ignore", True)
    np.savez_compressed(os.path.join(HERE, "golden_small.npz"), **g)

    surf = make_surface()
    print("wrote %d arrays, %d elements" % (len(g), len(surf)))


if __name__ == "__main__" and "--videofilters" in sys.argv:
    make_videofilters()
    sys.exit(0)
if __name__ == "__main__" and "--surface" in sys.argv:
    print("%d elements" % len(make_surface()))
    sys.exit(0)

if __name__ == "__main__":
    main()
