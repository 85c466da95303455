#!/usr/bin/env python
"""Break a kernel's executed-instruction count down by SASS region (from `ncu --page source --csv`).
usage: ncu_sass_segments.py report.ncu-rep kernel_name [dump_from dump_to]"""
import csv, io, subprocess, sys

rep, kern = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kern], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = next(r for r in rows if r and r[0] == "Address")
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows if r and r[0].startswith("0x")]
# several launches of the same kernel are concatenated: keep the first
first = data[0][0]
for j in range(1, len(data)):
    if data[j][0] == first:
        data = data[:j]; break
E = lambda r: int(r[ix["Instructions Executed"]])
tot = sum(E(r) for r in data)
print("total warp instructions", tot, "sass lines", len(data))
segs, start, acc, prev = [], 0, 0, None
for i, r in enumerate(data):
    n = E(r)
    if prev is not None and (n > prev * 1.3 + 1000 or n < prev / 1.3 - 1000):
        segs.append((start, i - 1, acc)); start, acc = i, 0
    acc += n; prev = n
segs.append((start, len(data) - 1, acc))
for a, b, c in segs:
    if c > tot * 0.005:
        print(f"{a:5d}-{b:5d} {c:12d} {100*c/tot:5.1f}%  exec/inst {E(data[a]):9d} lanes {data[a][ix['Avg. Threads Executed']]:>3}  {data[a][1].strip()[:48]}")
if len(sys.argv) > 4:
    for i in range(int(sys.argv[3]), int(sys.argv[4])):
        r = data[i]
        print(i, r[1].strip()[:72], E(r), r[ix["Avg. Threads Executed"]], r[ix["# Samples"]])
