#!/usr/bin/env python3
"""profiles/r02_sass_opcodes.md: for every kernel of libb200vf.so, how many TMA / bulk-copy / mbarrier / packed-fp32 /
vector-store / SIMD-in-word instructions its SASS holds (cuobjdump -sass of the built objects). Evidence for
"hand-written sm_100a": UTMALDG = cp.async.bulk.tensor (TMA tiles), UBLKCP = cp.async.bulk (1-D bulk copies),
SYNCS = mbarrier operations, FMUL2 / FFMA2 / FADD2 = packed f32x2, STG.128 / LDG.128 = 128-bit global accesses, VABSDIFF4 / IDP.4A =
four bytes per instruction. No HMMA / UTCMMA anywhere: byte work does not belong on tensor cores.
Usage: tools/sass_opcodes.py > profiles/r02_sass_opcodes.md   (after `make -C gst-plugins-bad_b200/csrc`)"""
import collections, glob, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COLS = ["UTMALDG", "UBLKCP", "SYNCS", "FMUL2", "FFMA2", "FADD2", "LDS.128", "STG.128", "LDG.128", "VABSDIFF4", "IDP.4A", "PRMT", "HMMA|UTCMMA"]
PATS = {"STG.128": r"STG\.[A-Z.]*128", "LDG.128": r"LDG\.[A-Z.]*128", "LDS.128": r"LDS\.[A-Z.]*128"}


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


rows = []
for obj in sorted(glob.glob(os.path.join(ROOT, "gst-plugins-bad_b200", "build", "*.o"))):
    if os.path.basename(obj).startswith("host_"):
        continue
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    name, counts, total = None, None, 0
    per = {}
    for line in sass.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            per[name] = [collections.Counter(), 0]
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and name:
            op = m.group(1)
            per[name][1] += 1
            for c in COLS:
                if re.match("(?:%s)" % PATS.get(c, c.replace(".", r"\.")), op):
                    per[name][0][c] += 1
    for k, (cnt, tot) in per.items():
        rows.append((os.path.basename(obj), k, cnt, tot))
dm = demangle([r[1] for r in rows])
print("# SASS opcode counts per kernel (sm_100a; `cuobjdump -sass gst-plugins-bad_b200/build/*.o`, tools/sass_opcodes.py)\n")
print(__doc__.split("Usage")[0].strip().split("\n", 1)[1].strip() + "\n")
print("| object | kernel | instrs | " + " | ".join(COLS) + " |")
print("|---|---|---|" + "---|" * len(COLS))
for obj, k, cnt, tot in rows:
    short = dm.get(k, k)
    short = re.sub(r"\(anonymous namespace\)::", "", short)
    short = re.sub(r"\(.*$", "", short)[:70]
    print("| %s | `%s` | %d | %s |" % (obj, short, tot, " | ".join(str(cnt[c]) if cnt[c] else "" for c in COLS)))
