"""CPU-only, world_size 2 over gloo: the host logic of the row-shard path.
Each rank owns the row block b200vf_shard_rows gives it, exchanges halo rows with its neighbour
(the transport the GPU path does with NCCL), processes [halo | shard | halo] and the concatenation of
the shards must equal the whole-frame result. The per-shard compute here is the oracle (this is a
test of partitioning / halo layout / edge rules, not of the kernels - those are the -m gpu tests)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIFT = {"bggr": "grbg", "grbg": "bggr", "gbrg": "rggb", "rggb": "gbrg"}   # pattern seen one row lower


def _worker(rank, world, port, q):
    sys.path.insert(0, os.path.join(ROOT, "gst-plugins-bad_b200"))
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import b200vf
    import oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    orc = oracle.best()
    rng = np.random.default_rng(0)                      # every rank draws the same global frames
    w, h = 64, 50
    r0, rows = b200vf.shard_rows(h, rank, world)

    def exchange(shard, halo):
        """send my first/last `halo` rows up/down, receive the neighbours' (what comm_halo_exchange does)"""
        top = bottom = None
        t = torch.from_numpy(np.ascontiguousarray(shard))
        reqs = []
        if rank > 0:
            top = torch.empty((halo,) + t.shape[1:], dtype=t.dtype)
            reqs += [dist.isend(t[:halo].contiguous(), rank - 1), dist.irecv(top, rank - 1)]
        if rank < world - 1:
            bottom = torch.empty((halo,) + t.shape[1:], dtype=t.dtype)
            reqs += [dist.isend(t[-halo:].contiguous(), rank + 1), dist.irecv(bottom, rank + 1)]
        [r.wait() for r in reqs]
        parts = ([top.numpy()] if top is not None else []) + [shard] + ([bottom.numpy()] if bottom is not None else [])
        return np.concatenate(parts, 0), (halo if top is not None else 0)

    res = {}
    # bayer2rgb: 1 halo row of the u8 mosaic
    mosaic = rng.integers(0, 256, (h, w), dtype=np.uint8)
    local, skip = exchange(mosaic[r0:r0 + rows], 1)
    for fmt in oracle.BAYER_FORMATS:
        lf = SHIFT[fmt] if skip else fmt               # local row 0 is global row r0-1 (odd) when a top halo exists
        out = orc.bayer2rgb(local, w, local.shape[0], lf, "RGBA")
        res["bayer_" + fmt] = out[skip:skip + rows]
    # dilate: 1 row below
    px = rng.integers(0, 2 ** 32, (h, w), dtype=np.uint32)
    local, skip = exchange(px[r0:r0 + rows], 1)
    res["dilate"] = orc.dilate(local, False)[skip:skip + rows]
    # gaussianblur sigma=2 (center 5): `center` halo rows, global truncation at the frame edges only
    fr = rng.integers(0, 256, (h, 4 * w), dtype=np.uint8)
    local, skip = exchange(fr[r0:r0 + rows], 5)
    res["gauss"] = orc.gaussblur(local, w, local.shape[0], 2.0, 0)[skip:skip + rows]
    # AYUV layout (p0 = 1), unpadded rows: the last channel of a row's last pixel lives in the next row, so the
    # shard needs one halo row more (b200vf_gaussblur_halo_rows); with only `center` rows the last shard row differs
    k, _ = b200vf.gauss_kernel(2.0)
    halo = b200vf.gaussblur_halo_rows(len(k), 1, 4 * w, w)
    assert halo == 6 and b200vf.gaussblur_halo_rows(len(k), 0, 4 * w, w) == 5 and b200vf.gaussblur_halo_rows(len(k), 1, 4 * w + 16, w) == 5
    local, skip = exchange(fr[r0:r0 + rows], halo)
    res["gauss_p1"] = orc.gaussblur(local, w, local.shape[0], 2.0, 1)[skip:skip + rows]
    q.put((rank, r0, rows, res))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_row_shards_reassemble_to_whole_frame(world):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle
    orc = oracle.best()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    got = sorted([q.get(timeout=120) for _ in range(world)])
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    rng = np.random.default_rng(0)
    w, h = 64, 50
    mosaic = rng.integers(0, 256, (h, w), dtype=np.uint8)
    px = rng.integers(0, 2 ** 32, (h, w), dtype=np.uint32)
    fr = rng.integers(0, 256, (h, 4 * w), dtype=np.uint8)
    cat = lambda k: np.concatenate([g[3][k] for g in got], 0)
    for fmt in oracle.BAYER_FORMATS:
        assert np.array_equal(cat("bayer_" + fmt), orc.bayer2rgb(mosaic, w, h, fmt, "RGBA")), fmt
    assert np.array_equal(cat("dilate"), orc.dilate(px, False))
    assert np.array_equal(cat("gauss"), orc.gaussblur(fr, w, h, 2.0, 0))
    assert np.array_equal(cat("gauss_p1"), orc.gaussblur(fr, w, h, 2.0, 1))
