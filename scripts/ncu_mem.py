#!/usr/bin/env python
"""Per memory instruction: L1 tag requests / shared wavefronts / L2 sectors (from the ncu source page),
listing the heaviest.  usage: python scripts/ncu_mem.py rep [queries]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; nq = float(sys.argv[2]) if len(sys.argv) > 2 else 1e6
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src))); hdr = rows[1]
ix = {k: hdr.index(k) for k in ["Source", "Instructions Executed", "L1 Tag Requests Global", "L1 Wavefronts Shared", "L2 Theoretical Sectors Global", "L2 Theoretical Sectors Local", "stall_long_sb", "stall_lg", "stall_short_sb", "stall_mio", "# Samples"]}
out = []
tot = {k: 0 for k in ix if k != "Source"}
for i, x in enumerate(rows[2:]):
    if len(x) <= max(ix.values()): continue
    v = {k: (float(x[j]) if k != "Source" and x[j] not in ("", "-") else 0) for k, j in ix.items() if k != "Source"}
    for k in tot: tot[k] += v[k]
    if v["L1 Tag Requests Global"] or v["L1 Wavefronts Shared"] or v["L2 Theoretical Sectors Local"]:
        out.append((i, x[ix["Source"]].strip(), v))
print("totals per query:", {k: round(t / nq, 1) for k, t in tot.items()})
for i, s, v in sorted(out, key=lambda t: -(t[2]["L1 Tag Requests Global"] + t[2]["L1 Wavefronts Shared"] + t[2]["L2 Theoretical Sectors Local"] / 4))[:45]:
    print(f"{i:5d} x{v['Instructions Executed']/nq:6.2f} tagreq/q={v['L1 Tag Requests Global']/nq:7.1f} shwf/q={v['L1 Wavefronts Shared']/nq:6.1f} l2sec/q={v['L2 Theoretical Sectors Global']/nq:7.1f} loc/q={v['L2 Theoretical Sectors Local']/nq:6.1f} longsb={v['stall_long_sb']:6.0f} lg={v['stall_lg']:5.0f} {s[:70]}")
