#!/usr/bin/env python3
"""Turns gpurun_out/<tag>_<kernel>.ncu-rep (+ launches_<tag>.csv) into the tracked summary under profiles/.
Usage: tools/ncu_summary.py <rep> <out.md> [launches.csv]"""
import csv, subprocess, sys, io, collections
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.avg.per_second", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
lines = ["# ncu summary of %s\n" % rep.split("/")[-1], "", "`ncu --set full --clock-control none --import-source on` (one launch; cold-cache, serialised).", ""]
for r in rows[2:]:
    lines.append("| metric | unit | value |\n|---|---|---|")
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            lines.append("| %s | %s | %s |" % (k, units[i], r[i]))
    stalls = [(hdr[i], r[i]) for i in range(len(hdr)) if hdr[i].startswith("smsp__average_warps_issue_stalled") and hdr[i].endswith("_per_issue_active.ratio")]
    stalls = sorted(stalls, key=lambda kv: -float(kv[1].replace(",", "") or 0))[:6]
    lines.append("\nTop warp stall reasons (warps per issue-active cycle):\n")
    for k, v in stalls:
        lines.append("* %s = %s" % (k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v))
    lines.append("")
if len(sys.argv) > 3:
    rr = [r for r in csv.reader(open(sys.argv[3])) if len(r) > 10]
    h = rr[0]; ki, vi, gi, bi = h.index("Kernel Name"), h.index("Metric Value"), h.index("Grid Size"), h.index("Block Size")
    t = collections.defaultdict(float); n = collections.Counter()
    for r in rr[1:]:
        try:
            t[r[ki]] += float(r[vi].replace(",", "")); n[r[ki]] += 1
        except ValueError:
            pass
    tot = sum(t.values())
    lines += ["## launch list of the same command (`--metrics gpu__time_duration.sum`, bench.py --steps 3 --warmup 3 --frames 24 --no-elements)", "",
              "| kernel | launches | total us | share |", "|---|---|---|---|"]
    for k in sorted(t, key=t.get, reverse=True)[:12]:
        lines.append("| `%s` | %d | %.1f | %.3f |" % (k[:100], n[k], t[k] / 1e3, t[k] / tot))
open(out, "w").write("\n".join(lines) + "\n")
print("wrote", out)
