"""4K timing of dilate / coloreffects / chromahold / exclusion / lut4 (run under gpurun; env knobs select variants)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gst-plugins-bad_b200"))
import torch, b200vf
ctx = b200vf.Context(0); side = torch.cuda.Stream(); torch.cuda.set_stream(side); st = side.cuda_stream
w, h, n = 3840, 2160, 16
a = torch.randint(0, 255, (n, h, 4 * w), dtype=torch.uint8, device="cuda"); b = torch.empty_like(a)
def timeit(f, it=30):
    for _ in range(5): f()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(side)
    for _ in range(it): f()
    e1.record(side); torch.cuda.synchronize(); return e0.elapsed_time(e1) / it * 1e-3
tab, ml = b200vf.coloreffects_table(2)
tab2, ml2 = b200vf.coloreffects_table(4)
lut = b200vf.lut_burn(175)
which = sys.argv[1:] or ["dilate", "coloreffects", "chromahold", "exclusion", "lut4"]
F = {"dilate": lambda: ctx.dilate(a, b, w, h, False, nframes=n, stream=st),
     "coloreffects": lambda: ctx.coloreffects_rgb(a, w, h, 4 * w, 4, (0, 1, 2), tab, ml, nframes=n, stream=st),
     "coloreffects_xpro": lambda: ctx.coloreffects_rgb(a, w, h, 4 * w, 4, (0, 1, 2), tab2, ml2, nframes=n, stream=st),
     "chromahold": lambda: ctx.chromahold(a, w, h, 4 * w, (0, 1, 2), (255, 0, 0), 30, nframes=n, stream=st),
     "exclusion": lambda: ctx.exclusion(a, b, n * w * h, 175, stream=st),
     "lut4": lambda: ctx.lut4(a, b, n * w * h, lut, stream=st)}
for k in which:
    t = timeit(F[k]); gbs = n * w * h * 8 / t / 1e9
    print("%-18s %9.1f fps  %7.1f GB/s  %.3f" % (k, n / t, gbs, gbs / 6570), flush=True)
