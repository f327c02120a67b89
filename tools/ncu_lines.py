#!/usr/bin/env python
"""Per-source-line executed-instruction / stall-sample breakdown of an .ncu-rep (needs -lineinfo and
--import-source on).  usage: ncu_lines.py report.ncu-rep units [top_n]"""
import collections, csv, io, subprocess, sys
rep, units = sys.argv[1], float(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
cur = ("?", 0, "")
fname = "?"
agg = collections.OrderedDict()
for r in csv.reader(io.StringIO(txt)):
    if len(r) >= 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if len(r) > 8 and r[0] == "" and r[2].startswith("0x"):
        try:
            ex, smp = int(r[7]), int(r[6])
        except ValueError:
            continue
        a = agg.setdefault(cur, [0, 0])
        a[0] += ex; a[1] += smp
    elif r and r[0].isdigit():
        cur = (fname, int(r[0]), ",".join(r[1:4])[:110])
tot = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values()) or 1
print("total warp-instr %d = %.2f lane-instr/unit" % (tot, tot * 32 / units))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%6.2f%% instr %5.1f%% stall  %s:%d  %s" % (100 * v[0] / tot, 100 * v[1] / ts, k[0], k[1], k[2]))
