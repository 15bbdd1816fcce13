"""Runs one element kernel a few times (for ncu -k regex:... captures of kernels bench.py --profile does not reach).
   python tools/run_one.py gaussblur|dilate|exclusion|chromahold|remap|remap_packed|lut4|coloreffects|fused|moments|direct|rgb2bayer|sad|videodiff|zebrastripe|smooth [4k|8k]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gst-plugins-bad_b200"))
import torch, b200vf
what = sys.argv[1]; size = sys.argv[2] if len(sys.argv) > 2 else "4k"
w, h = (3840, 2160) if size == "4k" else (7680, 4320)
ctx = b200vf.Context(0); side = torch.cuda.Stream(); torch.cuda.set_stream(side); st = side.cuda_stream
n = 2 if (what == "gaussblur" or size == "8k") else 8          # point ops: a batch larger than L2
a = torch.randint(0, 255, (n, h, 4 * w), dtype=torch.uint8, device="cuda"); b = torch.empty_like(a)
if what == "gaussblur":
    k, ks = b200vf.gauss_kernel(5.0)
    f = lambda: ctx.gaussblur(a, b, w, h, 4 * w, 1, k, ks, exact=True, nframes=n, stream=st)
elif what == "dilate":
    f = lambda: ctx.dilate(a, b, w, h, False, nframes=n, stream=st)
elif what == "exclusion":
    f = lambda: ctx.exclusion(a, b, n * w * h, 175, stream=st)
elif what == "chromahold":
    f = lambda: ctx.chromahold(a, w, h, 4 * w, (0, 1, 2), (255, 0, 0), 30, nframes=n, stream=st)
elif what == "lut4":
    lut = b200vf.lut_burn(175); f = lambda: ctx.lut4(a, b, n * w * h, lut, stream=st)
elif what == "remap":
    idx = torch.from_numpy(b200vf.gt_resolve_map(b200vf.gt_build_map("fisheye", w, h), w, h, 1)).cuda()
    f = lambda: ctx.remap(a, b, idx, w, h, 4, 4 * w, nframes=n, stream=st)
elif what == "remap_packed":
    idx = b200vf.gt_resolve_map(b200vf.gt_build_map("fisheye", w, h), w, h, 1)
    pk = torch.from_numpy(b200vf.gt_pack_index(idx, w, h)[0]).cuda()
    f = lambda: ctx.remap_packed(a, b, pk, w, h, nframes=n, stream=st)
elif what in ("sad", "videodiff", "zebrastripe", "smooth"):        # videofiltersbad / smooth on luma planes (1 B per sample)
    nl = 32 if size == "4k" else 8
    la = torch.randint(0, 255, (nl, h, w), dtype=torch.uint8, device="cuda"); lb = torch.randint(0, 255, (nl, h, w), dtype=torch.uint8, device="cuda")
    lo = torch.empty_like(la); sums = torch.zeros(nl, dtype=torch.int32, device="cuda")
    f = {"sad": lambda: ctx.sad_u8(la, lb, w, w, h, sums, nframes=nl, stream=st),
         "videodiff": lambda: ctx.videodiff_luma(la, lb, lo, w, w, h, nframes=nl, stream=st),
         "zebrastripe": lambda: ctx.zebrastripe(lb, 1, w, w, h, threshold=90, nframes=nl, stream=st),
         "smooth": lambda: ctx.smooth_plane(la, lo, w, w, h, nframes=2, stream=st)}[what]
elif what == "coloreffects":
    table, ml = b200vf.coloreffects_table(2)
    f = lambda: ctx.coloreffects_rgb(a, w, h, 4 * w, 4, (0, 1, 2), table, ml, nframes=n, stream=st)
elif what == "fused":
    src = torch.randint(0, 255, (8, h, w), dtype=torch.uint8, device="cuda"); dst = torch.empty((8, h, 4 * w), dtype=torch.uint8, device="cuda")
    table, ml = b200vf.coloreffects_table(2); sol = b200vf.lut_solarize()
    f = lambda: ctx.bayer2rgb_fused(src, w, dst, 4 * w, w, h, 0, (0, 1, 2), luma_table=table, lut=sol, nframes=8, stream=st)
elif what == "moments":
    nl = 32 if size == "4k" else 8
    la = torch.randint(0, 255, (nl, h, w), dtype=torch.uint8, device="cuda"); mom = torch.zeros(2 * nl, dtype=torch.int64, device="cuda")
    f = lambda: ctx.luma_moments(la, w, w, h, mom, nframes=nl, stream=st)
elif what == "rgb2bayer":
    mosaic = torch.empty((n, h, w), dtype=torch.uint8, device="cuda")
    f = lambda: ctx.rgb2bayer(a, 4 * w, mosaic, w, w, h, 0, nframes=n, stream=st)
elif what == "direct":
    src = torch.randint(0, 255, (8, h, w), dtype=torch.uint8, device="cuda"); dst = torch.empty((8, h, 4 * w), dtype=torch.uint8, device="cuda")
    ctx.set_variant("direct"); f = lambda: ctx.bayer2rgb(src, w, dst, 4 * w, w, h, 0, (0, 1, 2), nframes=8, stream=st)
for _ in range(3):
    f()
torch.cuda.synchronize()
print("ran", what, ctx.last_kernel())
