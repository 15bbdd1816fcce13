"""Frames/s of the other BASELINE.json configs on N GPUs (run under torchrun, one process per GPU):
  C3 gaussianblur sigma=5 3840x2160 (row shards + NCCL halo of 14 u8 rows),
  C4 fisheye 7680x4320 RGBA (output rows sharded, source all-gathered over NCCL),
  C5 chain bayer2rgb!coloreffects(sepia)!solarize 7680x4320 fused vs per-element launches (row shards + 1-row halo).
Weak scaling like bench.py: N x the frames, each rank owns 1/N of every frame. Device-timed (CUDA events), max over ranks."""
import ctypes, json, os, sys
import numpy as np
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gst-plugins-bad_b200"))
import b200vf


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = b200vf.Context(local)
    side = torch.cuda.Stream(); torch.cuda.set_stream(side); st = side.cuda_stream

    def bcast(b):
        t = torch.tensor(list(b), dtype=torch.uint8).cuda()
        if world > 1:
            dist.broadcast(t, 0)
        return bytes(t.cpu().tolist())
    comm = b200vf.Comm(ctx, rank, world, bcast) if world > 1 else None

    def timed(step, iters, warm=3):
        for _ in range(warm):
            step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(side)
        for _ in range(iters):
            step()
        b.record(side)
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b) / iters], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) * 1e-3

    out = {"n_gpus": world}
    # ---------------- C3 gaussianblur
    w, h, per_gpu = 3840, 2160, 4
    n = per_gpu * world
    k, ks = b200vf.gauss_kernel(5.0)
    halo = len(k) // 2 + 1
    r0, rows = b200vf.shard_rows(h, rank, world)
    rb = 4 * w
    hh = halo if world > 1 else 0
    fs = (rows + 2 * hh) * rb
    src = torch.randint(0, 255, (n, rows + 2 * hh, rb), dtype=torch.uint8, device="cuda")
    dst = torch.empty_like(src)

    def gstep():
        if comm:
            comm.halo_exchange(src, rb, rows, halo, fs, n, st)
        ctx.gaussblur(src.data_ptr() + hh * rb, dst.data_ptr() + hh * rb, w, rows, rb, 1, k, ks, nframes=n, frame_stride=fs,
                      row0=r0, rows=rows, full_height=h, stream=st)
    t = timed(gstep, 3)
    out["gaussblur_sigma5_4k"] = {"fps": n / t, "frames_per_step": n}
    del src, dst
    # ---------------- C4 fisheye 8K
    w, h, per_gpu = 7680, 4320, 2
    n = per_gpu * world
    r0, rows = b200vf.shard_rows(h, rank, world)
    rb = 4 * w
    full = torch.randint(0, 255, (n, h, rb), dtype=torch.uint8, device="cuda")
    dst = torch.empty((n, rows, rb), dtype=torch.uint8, device="cuda")
    idx = b200vf.gt_resolve_map(b200vf.gt_build_map("fisheye", w, h), w, h, 1)
    d_idx = torch.from_numpy(np.ascontiguousarray(idx[r0:r0 + rows])).cuda()
    bands = [b200vf.gt_index_row_range(idx[a:a + b], w) for (a, b) in [b200vf.shard_rows(h, r, world) for r in range(world)]]
    need_lo, need_hi = [b[0] for b in bands], [b[1] for b in bands]
    out["fisheye_bands"] = bands

    def fstep():
        if comm:
            comm.exchange_rows(full, rb, h, need_lo, need_hi, h * rb, n, st)
        for i in range(n):          # frames have different src/dst pitches here: one launch per frame
            ctx.remap(full[i], dst[i], d_idx, w, rows, 4, rb, stream=st)
    t = timed(fstep, 3)
    out["fisheye_8k"] = {"fps": n / t, "frames_per_step": n}
    del full, dst, d_idx
    # ---------------- C5 chain 8K
    per_gpu = 24
    n = per_gpu * world
    fsb = (rows + 2) * w
    src = torch.randint(0, 255, (n, rows + 2, w), dtype=torch.uint8, device="cuda")
    dst = torch.empty((n, rows, rb), dtype=torch.uint8, device="cuda")
    table, ml = b200vf.coloreffects_table(2)
    sol = b200vf.lut_solarize()

    def cfused():
        if comm:
            comm.halo_exchange(src, w, rows, 1, fsb, n, st)
        ctx.bayer2rgb_shard_fused(src.data_ptr() + w, w, dst, rb, w, h, r0, rows, 0, (0, 1, 2), luma_table=table, lut=sol,
                                  nframes=n, src_frame_stride=fsb, dst_frame_stride=rows * rb, stream=st)

    def cunfused():
        if comm:
            comm.halo_exchange(src, w, rows, 1, fsb, n, st)
        ctx.bayer2rgb_shard(src.data_ptr() + w, w, dst, rb, w, h, r0, rows, 0, (0, 1, 2), nframes=n, src_frame_stride=fsb,
                            dst_frame_stride=rows * rb, stream=st)
        ctx.coloreffects_rgb(dst, w, rows, rb, 4, (0, 1, 2), table, ml, nframes=n, stream=st)
        ctx.lut4(dst, dst, n * w * rows, sol, stream=st)
    t = timed(cfused, 5)
    out["chain_8k_fused"] = {"fps": n / t, "frames_per_step": n}
    t = timed(cunfused, 5)
    out["chain_8k_unfused"] = {"fps": n / t, "frames_per_step": n}
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
