#!/usr/bin/env python
"""Summarise an .ncu-rep: headline metrics + per-instruction execution / stall distribution.
usage: ncu_summary.py report.ncu-rep pixels_or_units [label]"""
import collections, csv, io, subprocess, sys

rep, units = sys.argv[1], float(sys.argv[2])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, unit, vals = rows[0], rows[1], rows[2]
d = dict(zip(hdr, vals)); u = dict(zip(hdr, unit))
keys = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__occupancy_limit_warps",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"]
for k in keys:
    if k in d:
        print("%-72s %s %s" % (k, d[k][:100], u[k]))
for k in hdr:
    if "issue_stalled" in k and k.endswith("per_issue_active.ratio"):
        try:
            if float(d[k]) >= 0.3:
                print("  stall %-60s %s" % (k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), d[k]))
        except ValueError:
            pass
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]
ia, isrc, iex, ismp = h.index("Address"), h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
data = []
for r in rows[2:]:
    if len(r) < len(h):
        break
    try:
        data.append((r[ia], r[isrc], int(r[iex]), int(r[ismp])))
    except ValueError:
        pass
tot = sum(x[2] for x in data); ts = sum(x[3] for x in data) or 1
print("static instr %d, executed warp-instr %d -> %.2f lane-instr per unit" % (len(data), tot, tot * 32 / units))
c = collections.Counter()
for x in data:
    toks = x[1].split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    c[op.split(".")[0]] += x[2]
print("per-unit by opcode:", [(k, round(v * 32 / units, 2)) for k, v in c.most_common(16)])
print("top stall sites:")
for x in sorted(data, key=lambda x: -x[3])[:10]:
    print("   %5.1f%%  %s" % (100 * x[3] / ts, x[1][:100]))
