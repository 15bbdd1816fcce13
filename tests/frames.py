"""Deterministic synthetic frames: seed-0 uniform random plus structured
frames (videotestsrc-like colour bars + ramp, constant, checkerboard, single hot pixel)."""
import numpy as np


def round_up_4(n):
    return (n + 3) & ~3


def random_u8(rng, h, stride):
    return rng.integers(0, 256, (h, stride), dtype=np.uint8)


def bars_rgbx(w, h, alpha=255):
    """SMPTE-like 7 colour bars over the top 2/3, luma ramp below; returns uint8 [h, 4*w] (R,G,B,x)."""
    cols = np.array([[191, 191, 191], [191, 191, 0], [0, 191, 191], [0, 191, 0], [191, 0, 191], [191, 0, 0],
                     [0, 0, 191]], np.uint8)
    img = np.zeros((h, w, 4), np.uint8)
    idx = np.minimum((np.arange(w) * 7) // max(w, 1), 6)
    img[: (2 * h) // 3, :, :3] = cols[idx][None, :, :]
    ramp = ((np.arange(w) * 255) // max(w - 1, 1)).astype(np.uint8)
    img[(2 * h) // 3:, :, :3] = ramp[None, :, None]
    img[:, :, 3] = alpha
    return img.reshape(h, 4 * w)


def mosaic_from_rgbx(rgbx, w, h, pattern="bggr"):
    """Sample an RGBx frame into a Bayer mosaic with pitch ROUND_UP_4(w)."""
    img = rgbx.reshape(h, w, 4)
    out = np.zeros((h, round_up_4(w)), np.uint8)
    # colour at (row parity, col parity) for each pattern: 0=R 1=G 2=B
    lay = {"bggr": [[2, 1], [1, 0]], "gbrg": [[1, 2], [0, 1]], "grbg": [[1, 0], [2, 1]], "rggb": [[0, 1], [1, 2]]}[pattern]
    for pr in (0, 1):
        for pc in (0, 1):
            out[pr::2, pc:w:2] = img[pr::2, pc::2, lay[pr][pc]]
    return out


def checker_u8(h, stride, a=0, b=255):
    y, x = np.mgrid[0:h, 0:stride]
    return np.where((x + y) & 1, b, a).astype(np.uint8)


def hot_pixel_u8(h, stride, y, x, v=255):
    f = np.zeros((h, stride), np.uint8)
    f[y, x] = v
    return f
