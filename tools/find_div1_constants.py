"""Developer tool (GPU): for divisors b within a few ulp of 1.0, find e with a / b == RN(a + a*e) for every fp32
dividend the blur produces (exhaustive, b200vf_gauss_selftest_div1). Prints table rows for kOneFmaDiv (gaussblur.cu)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gst-plugins-bad_b200"))
import b200vf
ctx = b200vf.Context(0)
lo, hi = (127 - 64) << 23, (127 + 13) << 23
for bits in range(0x3f800000 - 8, 0x3f800000 + 9):
    b = np.array([bits], np.uint32).view(np.float32)[0]
    e_exact = 1.0 / np.float64(b) - 1.0
    e0 = np.float32(e_exact)
    cands = [e0]
    for d in range(1, 4):
        up, dn = e0, e0
        for _ in range(d):
            up = np.nextafter(up, np.float32(np.inf)); dn = np.nextafter(dn, np.float32(-np.inf))
        cands += [up, dn]
    found = None
    for e in cands:
        if ctx.gauss_selftest_div1(b, e, 0, 1) == 0 and ctx.gauss_selftest_div1(b, e, lo, hi) == 0:
            found = e
            break
    print("  { 0x%08xu, 0x%08xu },   // b = %.9g, e = %s" % (bits, np.array([found if found is not None else 0], np.float32).view(np.uint32)[0],
                                                            b, "%.9g" % found if found is not None else "NONE"))
