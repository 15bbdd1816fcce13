"""Developer A/B of a gaussianblur environment knob in one process, interleaved (clocks drift under the power cap, so
single runs differ by a few per cent): python tools/time_gauss_ab.py B200VF_GAUSS_NO_AUX [frames per launch]"""
import os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gst-plugins-bad_b200"))
import torch
import b200vf

knob = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4
torch.cuda.set_device(0)
ctx = b200vf.Context(0)
side = torch.cuda.Stream()
torch.cuda.set_stream(side)
st = side.cuda_stream
w, h = 3840, 2160
k, ks = b200vf.gauss_kernel(5.0)
a = torch.randint(0, 256, (n, h, 4 * w), dtype=torch.uint8, device="cuda")
b = torch.empty_like(a)


def run(iters=10):
    for _ in range(2):
        ctx.gaussblur(a, b, w, h, 4 * w, 1, k, ks, exact=True, nframes=n, stream=st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(side)
    for _ in range(iters):
        ctx.gaussblur(a, b, w, h, 4 * w, 1, k, ks, exact=True, nframes=n, stream=st)
    e1.record(side)
    torch.cuda.synchronize()
    return n * iters / (e0.elapsed_time(e1) * 1e-3)


res = {"unset": [], "set": []}
for rep in range(8):
    for mode in ("unset", "set"):
        if mode == "set":
            os.environ[knob] = "1"
        else:
            os.environ.pop(knob, None)
        res[mode].append(run())
os.environ.pop(knob, None)
for mode in res:
    print("%s %s: median %.0f fps  (%s)" % (knob, mode, statistics.median(res[mode]), " ".join("%.0f" % v for v in res[mode])))
