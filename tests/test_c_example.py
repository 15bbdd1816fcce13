"""examples/chain_from_c.c: the drop-in boundary used from plain C (C99, -pedantic) - no Python, torch or GLib in the
process. CPU: it compiles and links against include/b200vf.h + libb200vf.so and, without a GPU, fails loudly at
b200vf_ctx_create (there is no CPU path). GPU: `bayer2rgb ! coloreffects preset=sepia ! solarize` through memories (one
fused launch, one transfer each way) and through host buffers gives the oracle's bytes."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "gst-plugins-bad_b200", "lib")


def build(tmp_path):
    exe = str(tmp_path / "chain_from_c")
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I" + os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "examples", "chain_from_c.c"), "-L" + LIBDIR, "-lb200vf", "-o", exe],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def run(exe, *args):
    env = dict(os.environ, LD_LIBRARY_PATH=LIBDIR + os.pathsep + os.environ.get("LD_LIBRARY_PATH", ""))
    return subprocess.run([exe] + [str(a) for a in args], capture_output=True, text=True, env=env, timeout=300)


def lcg_pattern(n):
    """the example's input: top byte of a 32-bit LCG (a = 1664525, c = 1013904223, seed 12345), vectorised by jumping"""
    out = np.empty(n, np.uint8)
    s = 12345
    block = 1 << 16
    # (a, c) of i steps: x -> a_i x + c_i
    a_i = np.empty(block, np.uint64)
    c_i = np.empty(block, np.uint64)
    a, c = 1, 0
    for i in range(block):
        a, c = (a * 1664525) & 0xffffffff, (c * 1664525 + 1013904223) & 0xffffffff
        a_i[i], c_i[i] = a, c
    for start in range(0, n, block):
        m = min(block, n - start)
        st = (a_i[:m] * np.uint64(s) + c_i[:m]) & np.uint64(0xffffffff)
        out[start:start + m] = (st >> np.uint64(24)).astype(np.uint8)
        s = int(st[m - 1])
    return out


def fnv1a64(data):
    h = 0xcbf29ce484222325
    for chunk in np.array_split(data, max(1, data.size // (1 << 16))):
        for b in chunk.tobytes():
            h = ((h ^ b) * 0x100000001b3) & 0xffffffffffffffff
    return h


def test_lcg_pattern_jump_equals_the_scalar_recurrence():
    s, want = 12345, []
    for _ in range(70000):
        s = (s * 1664525 + 1013904223) & 0xffffffff
        want.append(s >> 24)
    assert np.array_equal(lcg_pattern(70000), np.array(want, np.uint8))


def test_example_builds_as_c99_and_fails_loudly_without_a_gpu(tmp_path):
    exe = build(tmp_path)
    r = run(exe, 64, 48)
    import torch
    if torch.cuda.is_available():
        assert r.returncode == 0, r.stderr
    else:
        assert r.returncode == 1 and "b200vf_ctx_create" in r.stderr, (r.returncode, r.stderr)


@pytest.mark.gpu
@pytest.mark.parametrize("w,h", [(64, 48), (640, 480)])
def test_example_chain_equals_the_oracle(tmp_path, orc, w, h):
    exe = build(tmp_path)
    r = run(exe, w, h)
    assert r.returncode == 0, r.stderr
    fw, fh, hash_mem, hash_host, launches, h2d, d2h = r.stdout.split()
    assert (int(fw), int(fh)) == (w, h)
    src = lcg_pattern(w * h).reshape(h, w)
    rgb = orc.bayer2rgb(src, w, h, "bggr", "BGRx")
    want = orc.solarize(orc.coloreffects(rgb, w, h, "BGRx", "sepia").view(np.uint32)).view(np.uint8).reshape(-1)
    assert int(hash_mem, 16) == fnv1a64(want) and int(hash_host, 16) == fnv1a64(want)
    assert (int(launches), int(h2d), int(d2h)) == (1, 1, 1)
