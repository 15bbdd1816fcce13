"""What can this B200 sustain for the traffic MIX of bayer2rgb (1 B read : 4 B written)? Plain torch ops:
copy (1:1), fill (0:1), u8 -> int32 widening (1:4)."""
import torch, json, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
def t(fn, it=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / it * 1e-3
n = 1 << 30
src8 = torch.randint(0, 255, (n,), dtype=torch.uint8, device="cuda")
dst32 = torch.empty(n, dtype=torch.int32, device="cuda")
a = torch.empty(n, dtype=torch.int32, device="cuda"); 
print("copy 1:1   %.0f GB/s" % (8 * n / t(lambda: dst32.copy_(a)) / 1e9))
print("fill 0:1   %.0f GB/s" % (4 * n / t(lambda: dst32.fill_(7)) / 1e9))
print("widen 1:4  %.0f GB/s" % (5 * n / t(lambda: dst32.copy_(src8)) / 1e9))
print("read only  %.0f GB/s" % (4 * n / t(lambda: a.sum()) / 1e9))
