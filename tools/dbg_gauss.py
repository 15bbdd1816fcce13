import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gst-plugins-bad_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, b200vf, oracle
np.set_printoptions(linewidth=250)
ctx = b200vf.Context(0); orc = oracle.best()
rng = np.random.default_rng(0)
for (w, h, sigma, p0) in [(8, 8, 1.2, 0), (40, 30, 1.2, 0), (70, 66, 1.2, 0), (64, 64, 1.2, 0), (128, 128, 5, 0), (64,200,1.2,0), (200,64,1.2,0)]:
    fr = rng.integers(0, 256, (h, 4 * w), dtype=np.uint8)
    k, ks = b200vf.gauss_kernel(sigma)
    d_src = ctx.upload(fr); d_dst = ctx.alloc(fr.size)
    ctx.gaussblur(d_src, d_dst, w, h, 4 * w, p0, k, ks)
    got = ctx.download(d_dst, fr.size).reshape(fr.shape)
    want = orc.gaussblur(fr, w, h, sigma, p0)
    bad = np.argwhere(got != want)
    print(w, h, sigma, p0, "nbad", len(bad), "of", got.size)
    if len(bad):
        rows = np.unique(bad[:, 0]); cols = np.unique(bad[:, 1] // 4)
        print("  bad rows", rows[:40], "bad px cols", cols[:40])
        r, c = bad[0]
        print("  first", r, c, "got", got[r, c], "want", want[r, c], "maxdiff", np.abs(got.astype(int) - want).max())
