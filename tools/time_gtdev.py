"""Developer timing: int32 gather table of bulge at 8K built on the GPU vs on the host (+ upload)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gst-plugins-bad_b200"))
import numpy as np, torch, b200vf
ctx = b200vf.Context(0)
w, h = 7680, 4320
for el in ("bulge", "square", "mirror", "perspective"):
    b200vf.gt_build_index_device(ctx, el, w, h, {}, 1); ctx.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        d = b200vf.gt_build_index_device(ctx, el, w, h, {}, 1)
    ctx.synchronize()
    t_dev = (time.perf_counter() - t0) / 5
    t0 = time.perf_counter()
    idx = b200vf.gt_resolve_map(b200vf.gt_build_map(el, w, h), w, h, 1)
    up = ctx.upload(idx)
    t_host = time.perf_counter() - t0
    print("%s 8K: table on the GPU %.2f ms (incl. allocation) | host map + resolve + upload %.1f ms" % (el, t_dev * 1e3, t_host * 1e3))
