"""Developer timing: diffuse at 8K (per-pixel, per-frame draws + gather), frames/s and fraction of the measured HBM peak."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gst-plugins-bad_b200"))
import numpy as np, torch, b200vf
ctx = b200vf.Context(0)
w, h, n = 7680, 4320, 8
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6548.2) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6548.2
a = torch.randint(0, 256, (n, h, 4 * w), dtype=torch.uint8, device="cuda")
b = torch.empty_like(a)
s, c = b200vf.diffuse_tables(4.0)
ts = torch.cuda.Stream()                             # torch's events see only torch's streams; 0 would mean the context's own
torch.cuda.set_stream(ts)
st = ts.cuda_stream
for policy in (1, 0, 2):
    for _ in range(2):
        ctx.diffuse(a, b, w, h, 4, 4 * w, s, c, policy, 0, 1, 0, nframes=n, stream=st)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for i in range(5):
        ctx.diffuse(a, b, w, h, 4, 4 * w, s, c, policy, 0, 1, n * i, nframes=n, stream=st)
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 1e3 / (5 * n)
    print("diffuse 8K policy %d: %.1f us/frame, %.0f fps, %.3f of the HBM peak (8 B/px)" % (policy, t * 1e6, 1 / t, w * h * 8 / t / 1e9 / peak))
