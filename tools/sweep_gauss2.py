"""gaussblur timing over the library's tuning knobs in ONE process (run under gpurun).
   python tools/sweep_gauss2.py "" "GTH=96" "CTAS=148" "EDGEW=4" "PREPASS=1" ...   (knob = B200VF_GAUSS_<name>)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gst-plugins-bad_b200"))
import torch, b200vf
ctx = b200vf.Context(0); side = torch.cuda.Stream(); torch.cuda.set_stream(side); st = side.cuda_stream
w, h, n = 3840, 2160, 4
a = torch.randint(0, 255, (n, h, 4 * w), dtype=torch.uint8, device="cuda"); b = torch.empty_like(a)
cases = [(5.0, 1, True), (5.0, 0, True), (5.0, 1, False), (1.2, 1, True)]
def fps(sigma, p0, exact):
    k, ks = b200vf.gauss_kernel(sigma)
    f = lambda: ctx.gaussblur(a, b, w, h, 4 * w, p0, k, ks, exact=exact, nframes=n, stream=st)
    for _ in range(3): f()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(side)
    for _ in range(6): f()
    e1.record(side); torch.cuda.synchronize()
    return n * 6 / (e0.elapsed_time(e1) * 1e-3)
for _ in range(10): fps(5.0, 1, True)      # clocks
for cfg in (sys.argv[1:] or [""]):
    for k in list(os.environ):
        if k.startswith("B200VF_GAUSS_"): del os.environ[k]
    for kv in filter(None, cfg.split(",")):
        k, v = kv.split("="); os.environ["B200VF_GAUSS_" + k] = v
    print("%-28s" % (cfg or "default"), "  ".join("s%g/p%d/%s %7.1f" % (s, p, "ex" if e else "fma", fps(s, p, e)) for s, p, e in cases), flush=True)
