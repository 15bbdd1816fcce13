#!/bin/bash
# tuning sweeps (developer tool): bayer TMA CTAs per SM; gaussblur quick timing
for c in 2 3 4 5 6; do echo "== B200VF_TMA_CTAS_PER_SM=$c"; B200VF_TMA_CTAS_PER_SM=$c python tools/quick_bench.py --what bayer --variants auto --sizes 3840x2160 2>&1 | tail -1; done
python - <<'PY'
import sys, os
sys.path.insert(0, "gst-plugins-bad_b200")
import torch, b200vf
ctx = b200vf.Context(0); side = torch.cuda.Stream(); torch.cuda.set_stream(side); st = side.cuda_stream
w, h, n = 3840, 2160, 4
a = torch.randint(0, 255, (n, h, 4 * w), dtype=torch.uint8, device="cuda"); b = torch.empty_like(a)
for sigma in (5.0, 1.2, 20.0):
    k, ks = b200vf.gauss_kernel(sigma)
    for exact in (True, False):
        f = lambda: ctx.gaussblur(a, b, w, h, 4 * w, 1, k, ks, exact=exact, nframes=n, stream=st)
        for _ in range(2): f()
        torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(side)
        for _ in range(3): f()
        e1.record(side); torch.cuda.synchronize(); t = e0.elapsed_time(e1) / 3 * 1e-3
        print("gaussblur 4K sigma=%g taps=%d exact=%d: %.1f fps, %.2f T fp32 op/s" % (sigma, len(k), exact, n / t, 16 * len(k) * w * h * n / t / 1e12), flush=True)
PY
