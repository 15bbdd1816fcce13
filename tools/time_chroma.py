"""Developer A/B of chromahold (in place, data-dependent control flow): selects vs the literal branches, on frames
without spatial coherence (uniform random) and with it (colour bars + ramps), EVERY pass on fresh data - an in-place
element fed its own output sees grey frames from the second pass on, which take the shortest path."""
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
PEAK = 6548.2


def bars(n, h, w):
    """8 saturated colour bars across, a brightness ramp down, a little noise in the low bits: coherent but not constant"""
    x = torch.arange(w, device="cuda") * 8 // w
    cols = torch.tensor([[255, 255, 255], [255, 255, 0], [0, 255, 255], [0, 255, 0], [255, 0, 255], [255, 0, 0], [0, 0, 255], [40, 40, 40]],
                        device="cuda", dtype=torch.float32)
    ramp = (0.25 + 0.75 * torch.arange(h, device="cuda", dtype=torch.float32) / h)[:, None, None]
    img = (cols[x][None, :, :] * ramp).to(torch.uint8)                       # h x w x 3
    img = torch.cat([img, torch.full((h, w, 1), 255, dtype=torch.uint8, device="cuda")], dim=2)
    fr = img.reshape(1, h, 4 * w).repeat(n, 1, 1)
    return fr ^ torch.randint(0, 4, fr.shape, dtype=torch.uint8, device="cuda")


def timeit(fresh, work, w, h, n, iters=5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    total = 0.0
    for i in range(2 + iters):
        work.copy_(fresh)
        torch.cuda.synchronize()
        e0.record(side)
        ctx.chromahold(work, w, h, 4 * w, (0, 1, 2), (255, 0, 0), 30, nframes=n, stream=st)
        e1.record(side)
        torch.cuda.synchronize()
        if i >= 2:
            total += e0.elapsed_time(e1)
    return total / iters * 1e-3


for (w, h, n) in [(3840, 2160, 24), (7680, 4320, 6)]:
    data = {"random": torch.randint(0, 256, (n, h, 4 * w), dtype=torch.uint8, device="cuda"), "bars": bars(n, h, w)}
    work = torch.empty_like(data["random"])
    for name, fresh in data.items():
        for variant in ("selects", "branches"):
            if variant == "branches":
                os.environ["B200VF_CHROMA_BRANCHY"] = "1"
            else:
                os.environ.pop("B200VF_CHROMA_BRANCHY", None)
            t = timeit(fresh, work, w, h, n)
            print("chromahold %dx%d %-6s %-8s: %8.0f fps  %.3f of the HBM peak" % (w, h, name, variant, n / t, n * w * h * 8 / t / 1e9 / PEAK), flush=True)
os.environ.pop("B200VF_CHROMA_BRANCHY", None)
