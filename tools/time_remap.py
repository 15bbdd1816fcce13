"""fisheye 8K: int32 index table vs the step-coded table (run under gpurun)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gst-plugins-bad_b200"))
import torch, b200vf
ctx = b200vf.Context(0); side = torch.cuda.Stream(); torch.cuda.set_stream(side); st = side.cuda_stream
w, h, n = 7680, 4320, 4
a = torch.randint(0, 255, (n, h, 4 * w), dtype=torch.uint8, device="cuda"); b = torch.empty_like(a)
t0 = time.perf_counter(); idx = b200vf.gt_resolve_map(b200vf.gt_build_map("fisheye", w, h), w, h, 1); t1 = time.perf_counter()
packed, raw = b200vf.gt_pack_index(idx, w, h); t2 = time.perf_counter()
print("map+resolve %.2fs  pack %.2fs  raw groups %d  packed %.2f B/px" % (t1 - t0, t2 - t1, raw, packed.size / (w * h)))
d_idx = torch.from_numpy(idx).cuda(); d_p = torch.from_numpy(packed).cuda()
def timeit(f, it=60):
    for _ in range(10): f()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(side)
    for _ in range(it): f()
    e1.record(side); torch.cuda.synchronize(); return e0.elapsed_time(e1) / it * 1e-3
for name, f in (("int32 table", lambda: ctx.remap(a, b, d_idx, w, h, 4, 4 * w, nframes=n, stream=st)),
                ("packed table", lambda: ctx.remap_packed(a, b, d_p, w, h, nframes=n, stream=st))):
    t = timeit(f); print("%-13s %8.1f fps  %6.1f GB/s algorithmic (8 B/px)" % (name, n / t, n * w * h * 8 / t / 1e9), flush=True)
