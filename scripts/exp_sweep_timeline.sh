#!/bin/bash
# launch sequence (ncu launch list, chronological) around the 256-obstacle add sweeps of the bench
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-c1 --no-c4 --no-c5"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/sweep_launches.csv $CMD > /dev/null 2>&1
python - <<'P'
import csv
rows = list(csv.reader(open("gpurun_out/sweep_launches.csv", errors="replace")))
hdr = next(r for r in rows if "Kernel Name" in r)
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
seq = []
for r in rows[rows.index(hdr) + 1:]:
    try: seq.append((r[ki].split("(")[0][-60:], float(r[vi].replace(",", "")) / 1000.0))
    except Exception: pass
idx = [i for i, (n, v) in enumerate(seq) if "ig_test_kernel" in n and "Sweep" in n and v > 50]
for i in idx[-2:]:
    print("----")
    for n, v in seq[max(0, i - 8): i + 5]: print(f"{v:9.1f} us  {n}")
idx = [i for i, (n, v) in enumerate(seq) if "ig_test_kernel" in n and "Check" in n]
for i in idx[-1:]:
    print("---- check")
    for n, v in seq[max(0, i - 8): i + 5]: print(f"{v:9.1f} us  {n}")
P
