#!/bin/bash
# gaussblur 4K timing for a few tile heights / knobs (run under gpurun). Usage: tools/sweep_gauss.sh [gth ...]
GTHS=${@:-96 80 64}
run() { python - <<'PY'
import sys; sys.path.insert(0, "gst-plugins-bad_b200")
import torch, b200vf
ctx = b200vf.Context(0); side = torch.cuda.Stream(); torch.cuda.set_stream(side); st = side.cuda_stream
w, h, n = 3840, 2160, 4
a = torch.randint(0, 255, (n, h, 4 * w), dtype=torch.uint8, device="cuda"); b = torch.empty_like(a)
for sigma in (5.0, 1.2):
    k, ks = b200vf.gauss_kernel(sigma)
    for p0 in (1, 0):
        for exact in (True, False):
            f = lambda: ctx.gaussblur(a, b, w, h, 4 * w, p0, k, ks, exact=exact, nframes=n, stream=st)
            for _ in range(3): f()
            torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(side)
            for _ in range(5): f()
            e1.record(side); torch.cuda.synchronize(); t = e0.elapsed_time(e1) / 5 * 1e-3
            print("  sigma=%g p0=%d exact=%d: %.1f fps" % (sigma, p0, exact, n / t), flush=True)
PY
}
echo "== warm-up process (clocks)"; B200VF_GAUSS_GTH=96 run > /dev/null
for g in $GTHS; do echo "== B200VF_GAUSS_GTH=$g"; B200VF_GAUSS_GTH=$g run; done
echo "== GTH=96, __fdiv_rn instead of the reciprocal division"; B200VF_GAUSS_GTH=96 B200VF_GAUSS_NO_FASTDIV=1 run
