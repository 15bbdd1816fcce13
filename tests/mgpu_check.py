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
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    b200vf.lib.b200vf_comm_destroy(comm)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
