"""Run under torchrun (one process per GPU): row-sharded bayer2rgb with the NCCL halo exchange of
b200vf_comm_*, gathered on rank 0 and compared bit-exactly with the whole-frame oracle.
   python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tests/mgpu_check.py"""
import ctypes
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gst-plugins-bad_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import b200vf  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = b200vf.Context(local)
    side = torch.cuda.Stream()
    torch.cuda.set_stream(side)
    st = side.cuda_stream
    idt = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        buf = (ctypes.c_uint8 * 128)()
        b200vf.check(b200vf.lib.b200vf_comm_unique_id(buf))
        idt = torch.tensor(list(buf), dtype=torch.uint8)
    idt = idt.cuda()
    dist.broadcast(idt, 0)
    idb = (ctypes.c_uint8 * 128)(*idt.cpu().tolist())
    comm = ctypes.c_void_p()
    b200vf.check(b200vf.lib.b200vf_comm_create(ctx.h, idb, rank, world, ctypes.byref(comm)))

    ok = True
    for (w, h, n, variant) in [(512, 192, 3, "auto"), (512, 192, 3, "direct"), (3840, 2160, 2, "auto")]:
        ctx.set_variant(variant)
        rng = np.random.default_rng(5)                       # same global frames on every rank
        frames = rng.integers(0, 256, (n, h, w), dtype=np.uint8)
        r0, rows = b200vf.shard_rows(h, rank, world)
        fs = (rows + 2) * w
        local_buf = np.zeros((n, rows + 2, w), np.uint8)     # [halo | shard | halo], halos filled by the exchange
        local_buf[:, 1:rows + 1] = frames[:, r0:r0 + rows]
        src = torch.from_numpy(local_buf).cuda()
        dst = torch.zeros((n, rows, 4 * w), dtype=torch.uint8, device="cuda")
        b200vf.check(b200vf.lib.b200vf_comm_halo_exchange(comm, src.data_ptr(), w, rows, 1, fs, n, st))
        ctx.bayer2rgb_shard(src.data_ptr() + w, w, dst, 4 * w, w, h, r0, rows, 0, (0, 1, 2), nframes=n,
                            src_frame_stride=fs, dst_frame_stride=rows * 4 * w, stream=st)
        torch.cuda.synchronize()
        kernel = ctx.last_kernel()
        sizes = [b200vf.shard_rows(h, r, world)[1] for r in range(world)]
        if rank == 0:
            parts = [dst.cpu().numpy()]
            for r in range(1, world):
                t = torch.empty((n, sizes[r], 4 * w), dtype=torch.uint8, device="cuda")
                dist.recv(t, r)
                parts.append(t.cpu().numpy())
            got = np.concatenate(parts, 1)
            import oracle
            orc = oracle.best()
            for i in range(n):
                good = np.array_equal(got[i], orc.bayer2rgb(frames[i], w, h, "bggr", "RGBA"))
                ok = ok and good
                print("mgpu %dx%d frame %d world=%d kernel=%s: %s" % (w, h, i, world, kernel, "OK" if good else "MISMATCH"), flush=True)
        else:
            dist.send(dst, 0)
    ctx.set_variant("auto")
    import oracle
    orc = oracle.best() if rank == 0 else None

    def gather_rows(t, h, n, rowbytes):
        """rank 0 receives every rank's [n, rows_r, rowbytes] shard and stacks them along rows"""
        sizes = [b200vf.shard_rows(h, r, world)[1] for r in range(world)]
        if rank == 0:
            parts = [t.cpu().numpy()]
            for r in range(1, world):
                x = torch.empty((n, sizes[r], rowbytes), dtype=torch.uint8, device="cuda")
                dist.recv(x, r)
                parts.append(x.cpu().numpy())
            return np.concatenate(parts, 1)
        dist.send(t.contiguous(), 0)
        return None

    # ---- gaussianblur sigma=5 (13-row halo, +1 above because p0 = 1), AYUV-faithful p0 = 1
    w, h, n, sigma, p0 = 256, 400, 2, 5.0, 1
    rng = np.random.default_rng(11)
    fr = rng.integers(0, 256, (n, h, 4 * w), dtype=np.uint8)
    k, ks = b200vf.gauss_kernel(sigma)
    halo = len(k) // 2 + 1
    r0, rows = b200vf.shard_rows(h, rank, world)
    rb = 4 * w
    fs = (rows + 2 * halo) * rb
    loc = np.zeros((n, rows + 2 * halo, rb), np.uint8)
    loc[:, halo:halo + rows] = fr[:, r0:r0 + rows]
    src = torch.from_numpy(loc).cuda()
    dst = torch.zeros((n, rows, rb), dtype=torch.uint8, device="cuda")
    b200vf.check(b200vf.lib.b200vf_comm_halo_exchange(comm, src.data_ptr(), rb, rows, halo, fs, n, st))
    # dst has its own frame pitch: run frame by frame (src/dst pitches differ)
    for i in range(n):
        ctx.gaussblur(src.data_ptr() + i * fs + halo * rb, dst.data_ptr() + i * rows * rb, w, rows, rb, p0, k, ks,
                      nframes=1, row0=r0, rows=rows, full_height=h, stream=st)
    torch.cuda.synchronize()
    got = gather_rows(dst, h, n, rb)
    if rank == 0:
        for i in range(n):
            want = orc.gaussblur(fr[i], w, h, sigma, p0)
            # byte 0 of the frame is never produced by the blur (the element copies it); shards leave it 0
            g = got[i].copy(); g[0, :p0] = want[0, :p0]
            good = np.array_equal(g, want)
            ok = ok and good
            print("mgpu gaussblur %dx%d sigma=%g frame %d world=%d: %s" % (w, h, sigma, i, world, "OK" if good else "MISMATCH %d" % (g != want).sum()), flush=True)

    # ---- dilate: 1 row below
    px = rng.integers(0, 2 ** 32, (h, w), dtype=np.uint32)
    loc = np.zeros((rows + 2, w), np.uint32)
    loc[1:rows + 1] = px[r0:r0 + rows]
    src = torch.from_numpy(loc.view(np.uint8).reshape(1, rows + 2, rb)).cuda()
    dst = torch.zeros((1, rows, rb), dtype=torch.uint8, device="cuda")
    b200vf.check(b200vf.lib.b200vf_comm_halo_exchange(comm, src.data_ptr(), rb, rows, 1, (rows + 2) * rb, 1, st))
    below = src.data_ptr() + (rows + 1) * rb if rank < world - 1 else None
    ctx.dilate(src.data_ptr() + rb, dst, w, rows, False, nframes=1, below=below, stream=st)
    torch.cuda.synchronize()
    got = gather_rows(dst, h, 1, rb)
    if rank == 0:
        good = np.array_equal(got[0].view(np.uint32).reshape(h, w), orc.dilate(px, False))
        ok = ok and good
        print("mgpu dilate %dx%d world=%d: %s" % (w, h, world, "OK" if good else "MISMATCH"), flush=True)

    # ---- fisheye remap: output rows sharded, source all-gathered (any source row may be read)
    full = torch.zeros((h, rb), dtype=torch.uint8, device="cuda")
    frame = rng.integers(0, 256, (h, rb), dtype=np.uint8)
    full[r0:r0 + rows] = torch.from_numpy(frame[r0:r0 + rows]).cuda()
    torch.cuda.synchronize()
    m = b200vf.gt_build_map("fisheye", w, h)
    idx = b200vf.gt_resolve_map(m, w, h, 1)
    # banded exchange: only the source rows each output shard reads travel (all-gather is the fallback)
    bands = [b200vf.gt_index_row_range(idx[a:a + b], w) for (a, b) in [b200vf.shard_rows(h, r, world) for r in range(world)]]
    lo = np.array([b[0] for b in bands], np.int32); hi = np.array([b[1] for b in bands], np.int32)
    b200vf.check(b200vf.lib.b200vf_comm_exchange_rows(comm, full.data_ptr(), rb, h, lo.ctypes.data_as(ctypes.c_void_p),
                                                      hi.ctypes.data_as(ctypes.c_void_p), 0, 1, st))
    d_idx = torch.from_numpy(np.ascontiguousarray(idx[r0:r0 + rows])).cuda()
    dst = torch.zeros((1, rows, rb), dtype=torch.uint8, device="cuda")
    ctx.remap(full, dst, d_idx, w, rows, 4, rb, stream=st)           # a shard of output rows: `height` = rows of this shard
    torch.cuda.synchronize()
    got = gather_rows(dst, h, 1, rb)
    if rank == 0:
        good = np.array_equal(got[0], orc.remap(frame, m, w, h, 4, "clamp", False))
        ok = ok and good
        print("mgpu fisheye %dx%d world=%d: %s" % (w, h, world, "OK" if good else "MISMATCH"), flush=True)

    # ---- fused chain bayer2rgb ! coloreffects(sepia) ! solarize on row shards
    mosaic = rng.integers(0, 256, (h, w), dtype=np.uint8)
    loc = np.zeros((1, rows + 2, w), np.uint8)
    loc[0, 1:rows + 1] = mosaic[r0:r0 + rows]
    src = torch.from_numpy(loc).cuda()
    dst = torch.zeros((1, rows, rb), dtype=torch.uint8, device="cuda")
    b200vf.check(b200vf.lib.b200vf_comm_halo_exchange(comm, src.data_ptr(), w, rows, 1, (rows + 2) * w, 1, st))
    table, ml = b200vf.coloreffects_table(2)
    ctx.bayer2rgb_shard_fused(src.data_ptr() + w, w, dst, rb, w, h, r0, rows, 0, (2, 1, 0), luma_table=table,
                              lut=b200vf.lut_solarize(), stream=st)
    torch.cuda.synchronize()
    got = gather_rows(dst, h, 1, rb)
    if rank == 0:
        rgb = orc.bayer2rgb(mosaic, w, h, "bggr", "BGRx")
        want = orc.solarize(orc.coloreffects(rgb, w, h, "BGRx", "sepia").view(np.uint32)).view(np.uint8).reshape(h, rb)
        good = np.array_equal(got[0], want)
        ok = ok and good
        print("mgpu fused chain %dx%d world=%d kernel=%s: %s" % (w, h, world, ctx.last_kernel(), "OK" if good else "MISMATCH"), flush=True)

    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    b200vf.lib.b200vf_comm_destroy(comm)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
