"""Developer timing: fisheye 8K remap, one frame per launch over a ring of frames (index table from HBM every frame) and
batched; B200VF_REMAP_BLOCKS_PER_SM sweeps the grid."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gst-plugins-bad_b200"))
import torch, b200vf
torch.cuda.set_device(0); ctx = b200vf.Context(0); side = torch.cuda.Stream(); torch.cuda.set_stream(side); st = side.cuda_stream
w, h, n = 7680, 4320, 11
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6548.0
a = torch.randint(0, 255, (n, h, 4 * w), dtype=torch.uint8, device="cuda"); b = torch.empty_like(a)
idx = torch.from_numpy(b200vf.gt_resolve_map(b200vf.gt_build_map("fisheye", w, h), w, h, 1)).cuda()
def timeit(fn, iters):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(side)
    for _ in range(iters): fn()
    e1.record(side); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3
cnt = [0]
def one():
    i = cnt[0] % n; cnt[0] += 1
    ctx.remap(a[i], b[i], idx, w, h, 4, 4 * w, nframes=1, stream=st)
t1 = timeit(one, 2 * n)
tb = timeit(lambda: ctx.remap(a, b, idx, w, h, 4, 4 * w, nframes=n, stream=st), 5) / n
px = w * h
print("blocks/SM %s: single frame %.1f us (%.3f of peak on 8 B/px, %.3f on 12 B/px) | batch of %d: %.1f us/frame (%.3f on 8 B/px)" % (
    os.environ.get("B200VF_REMAP_BLOCKS_PER_SM", "auto"), t1 * 1e6, px * 8 / t1 / 1e9 / peak, px * 12 / t1 / 1e9 / peak, n, tb * 1e6, px * 8 / tb / 1e9 / peak))
