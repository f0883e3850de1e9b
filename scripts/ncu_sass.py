#!/usr/bin/env python
"""Per-SASS-instruction execution counts of an .ncu-rep, grouped into straight-line segments of equal
execution count (= loop bodies / phases).  usage: python scripts/ncu_sass.py rep [dump.txt] [queries]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
dump = sys.argv[2] if len(sys.argv) > 2 else None
nq = float(sys.argv[3]) if len(sys.argv) > 3 else 1e6
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
iS, iE, iN = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
ins = [(x[iS].strip(), int(x[iE]), int(x[iN])) for x in rows[2:] if len(x) > iE]
tot = sum(e for _, e, _ in ins); ts = sum(n for _, _, n in ins)
print(f"total warp-instructions {tot}  = {tot/nq:.0f} per query; samples {ts}")
if dump:
    open(dump, "w").write("\n".join(f"{i:5d} {e:12d} {n:6d} {s}" for i, (s, e, n) in enumerate(ins)))
seg = []; cur = [0]
for i in range(1, len(ins)):
    a, b = ins[cur[-1]][1], ins[i][1]
    if (a == 0 and b == 0) or (a > 0 and b > 0 and 0.8 < b / a < 1.25): cur.append(i)
    else: seg.append(cur); cur = [i]
seg.append(cur)
for s in seg:
    e = sum(ins[i][1] for i in s); n = sum(ins[i][2] for i in s)
    if e / tot > 0.004 or n / ts > 0.01:
        ops = {}
        for i in s:
            t = ins[i][0].split()
            op = t[1] if t[0].startswith("@") else t[0]
            ops[op] = ops.get(op, 0) + 1
        top = " ".join(f"{k}:{v}" for k, v in sorted(ops.items(), key=lambda x: -x[1])[:7])
        print(f"{s[0]:5d}-{s[-1]:5d} n={len(s):4d} x{ins[s[0]][1]/nq:7.2f}/q  inst/q={e/nq:7.1f} ({e/tot*100:4.1f}%) stall {n/ts*100:4.1f}%  {top}")
