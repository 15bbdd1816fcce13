"""Developer timing of smooth (luma plane, default tolerance 8 / filter-size 3); B200VF_SMOOTH_SCALAR=1 times round 1's kernel."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gst-plugins-bad_b200"))
import torch, b200vf
ctx = b200vf.Context(0); side = torch.cuda.Stream(); torch.cuda.set_stream(side); st = side.cuda_stream
for (w, h) in [(3840, 2160), (7680, 4320)]:
    n = 8
    a = torch.randint(0, 255, (n, h, w), dtype=torch.uint8, device="cuda"); b = torch.empty_like(a)
    for tol, fs in ((8, 3), (200, 3), (8, 1), (8, 6)):
        f = lambda: ctx.smooth_plane(a, b, w, w, h, tolerance=tol, filtersize=fs, nframes=n, stream=st)
        for _ in range(2): f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(side)
        for _ in range(3): f()
        e1.record(side); torch.cuda.synchronize()
        t = e0.elapsed_time(e1) / 3 * 1e-3
        print("%dx%d tolerance %d filter-size %d: %s %.0f fps" % (w, h, tol, fs, ctx.last_kernel(), n / t))
