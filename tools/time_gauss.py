"""Developer timing of gaussianblur (not the contract bench): frames/s and fraction of the nominal fp32 roofline
(16 T lane-operations per pixel) for several sigmas / layouts / batch sizes; B200VF_GAUSS_NO_STREAM=1
in the environment times the general kernel."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gst-plugins-bad_b200"))
import torch
import b200vf

torch.cuda.set_device(0)
ctx = b200vf.Context(0)
side = torch.cuda.Stream()
torch.cuda.set_stream(side)
st = side.cuda_stream
FP32 = 148 * 128 * 1.965e9


def timeit(fn, iters=5):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(side)
    for _ in range(iters):
        fn()
    b.record(side)
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e-3


cases = [(3840, 2160, 5.0, 1, 1), (3840, 2160, 5.0, 1, 2), (3840, 2160, 5.0, 1, 8), (3840, 2160, 5.0, 1, 16), (3840, 2160, 5.0, 1, 32),
         (3840, 2160, 5.0, 0, 16), (7680, 4320, 5.0, 1, 1), (7680, 4320, 5.0, 1, 8),
         (3840, 2160, 5.0, 1, 4), (3840, 2160, 5.0, 2, 4), (3840, 2160, 5.0, 0, 4), (3840, 2160, 5.0, 1, 1), (7680, 4320, 5.0, 1, 4),
         (3840, 2160, 1.0, 1, 4), (3840, 2160, 1.2, 1, 4), (3840, 2160, 2.0, 1, 4), (3840, 2160, 2.7, 1, 4), (3840, 2160, 3.3, 1, 4),
         (3840, 2160, 4.3, 1, 4), (1920, 1080, 5.0, 1, 1)]
for (w, h, sigma, p0, n) in cases:
    k, ks = b200vf.gauss_kernel(sigma)
    a = torch.randint(0, 256, (n, h, 4 * w), dtype=torch.uint8, device="cuda")
    b = torch.empty_like(a)
    t = timeit(lambda: ctx.gaussblur(a, b, w, h, 4 * w, p0, k, ks, exact=True, nframes=n, stream=st))
    flops = 16 * len(k) * w * h * n
    print("%dx%d sigma %.1f (%d taps) p0 %d n %d: %8.1f fps  %.3f of fp32 roofline  [%s]" % (
        w, h, sigma, len(k), p0, n, n / t, flops / t / FP32, ctx.last_kernel()), flush=True)
    del a, b
