"""Developer timing: int32 gather tables at 8K built on the GPU (exact / certified + host patches) vs on the host (+ upload)."""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gst-plugins-bad_b200"))
import numpy as np, torch, b200vf
ctx = b200vf.Context(0)
w, h = 7680, 4320
buf = b200vf.DeviceBuffer(ctx, w * h * 4)
names, vals = (C.c_char_p * 1)(), (C.c_double * 1)()
for el in (sys.argv[1:] or ("bulge", "square", "mirror", "perspective", "marble", "fisheye", "circle", "kaleidoscope", "pinch", "sphere", "twirl", "waterripple")):
    def build():
        b200vf.check(b200vf.lib.b200vf_gt_build_index_device(ctx.h, el.encode(), w, h, names, vals, 0, 1, buf.ptr, None))
    build(); ctx.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        build()
    ctx.synchronize()
    t_dev = (time.perf_counter() - t0) / 5
    t0 = time.perf_counter()
    idx = b200vf.gt_resolve_map(b200vf.gt_build_map(el, w, h), w, h, 1)
    up = ctx.upload(idx)
    t_host = time.perf_counter() - t0
    del up
    print("%-12s 8K: table on the GPU %6.2f ms (%6d entries from the host) | host map + resolve + upload %6.1f ms"
          % (el, t_dev * 1e3, b200vf.gt_device_last_uncertain(), t_host * 1e3))
