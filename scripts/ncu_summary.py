#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU): key counters + per-source-line instruction / stall shares.
usage: python scripts/ncu_summary.py gpurun_out/prof.ncu-rep [top_lines]"""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "smsp__inst_executed.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size"]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
    print("== kernel:", name[:100])
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print(f"  {w:70s} {r[i]:>16s} {units[i]}")
    for i, h in enumerate(hdr):
        if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
            try:
                v = float(r[i])
            except ValueError:
                continue
            if v >= 0.1:
                print(f"  stall {h.split('issue_stalled_')[1].split('_per_issue')[0]:28s} {v:8.3f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True,
                     text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
cur = None
hdr = None
agg = []
ops = collections.Counter()
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        ia = hdr.index("Instructions Executed")
        ist = hdr.index("Warp Stall Sampling (All Samples)")
        continue
    if hdr is None or r[0] in ("Function Name",):
        continue
    try:
        if r[0].isdigit() and r[2] == "-":
            agg.append((cur, int(r[0]), r[1].strip(), int(r[ia]), int(r[ist])))
        elif r[0] == "-" or not r[0].isdigit():
            m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[3])
            if m:
                ops[m.group(2)] += int(r[ia])
    except (ValueError, IndexError):
        pass
tot = sum(a[3] for a in agg) or 1
ts = sum(a[4] for a in agg) or 1
print(f"== source lines (total warp-instructions {tot}, stall samples {ts})")
for a in sorted(agg, key=lambda a: -a[3])[:top]:
    print(f"  {a[0]:18s}:{a[1]:4d} inst {a[3] / tot * 100:5.2f}%  stall {a[4] / ts * 100:5.2f}%  {a[2][:80]}")
ot = sum(ops.values()) or 1
print("== opcodes:", ", ".join(f"{o} {c / ot * 100:.1f}%" for o, c in ops.most_common(18)))
