"""Developer micro-benchmark (not the contract bench): device-resident frames, CUDA-event timing
on the launching stream, frames larger than L2 in total so nothing is served from cache."""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gst-plugins-bad_b200"))
import torch
import b200vf

PEAK = 6570.0
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def timeit(fn, iters=20, warm=3):
    s = torch.cuda.current_stream()
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(iters):
        fn()
    e1.record(s)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--what", default="bayer,lut")
    ap.add_argument("--sizes", default="3840x2160,7680x4320")
    ap.add_argument("--variants", default="direct,auto")
    a = ap.parse_args()
    torch.cuda.set_device(0)
    ctx = b200vf.Context(0)
    side = torch.cuda.Stream()          # a real (non-NULL) stream: NULL means "the context's stream" in the C-ABI
    torch.cuda.set_stream(side)
    st = side.cuda_stream
    for size in a.sizes.split(","):
        w, h = map(int, size.split("x"))
        n = max(4, int(1.5e9 // (w * h * 5)))      # >= 1.5 GB per pass: far beyond the 126 MB L2
        if "bayer" in a.what:
            src = torch.randint(0, 256, (n, h, w), dtype=torch.uint8, device="cuda")
            dst = torch.empty((n, h, w * 4), dtype=torch.uint8, device="cuda")
            for var in a.variants.split(","):
                ctx.set_variant(var)
                try:
                    t = timeit(lambda: ctx.bayer2rgb(src, w, dst, 4 * w, w, h, 0, (0, 1, 2), nframes=n, stream=st))
                except b200vf.B200vfError as e:
                    print("bayer", size, var, "ERR", e); continue
                gbs = n * w * h * 5 / t / 1e9
                print("bayer2rgb %s %-6s kernel=%-18s %8.1f fps  %7.1f GB/s  %.3f of %.0f" % (
                    size, var, ctx.last_kernel(), n / t, gbs, gbs / PEAK, PEAK), flush=True)
            ctx.set_variant("auto")
            del src, dst
        if "lut" in a.what:
            n2 = max(2, int(1.5e9 // (w * h * 8)))
            src = torch.randint(0, 2 ** 31, (n2, h, w), dtype=torch.int32, device="cuda")
            dst = torch.empty_like(src)
            lut = b200vf.lut_burn(175)
            t = timeit(lambda: ctx.lut4(src, dst, n2 * w * h, lut, stream=st))
            gbs = n2 * w * h * 8 / t / 1e9
            print("lut4(burn) %s %8.1f fps  %7.1f GB/s  %.3f of %.0f" % (size, n2 / t, gbs, gbs / PEAK, PEAK), flush=True)
            # plain copy for reference
            t = timeit(lambda: dst.copy_(src))
            gbs = n2 * w * h * 8 / t / 1e9
            print("torch copy  %s %7.1f GB/s  %.3f" % (size, gbs, gbs / PEAK), flush=True)
            del src, dst


if __name__ == "__main__":
    main()
