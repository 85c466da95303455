#!/usr/bin/env python
"""Warp-stall samples of one kernel aggregated per enclosing source function (and per source line).
usage: ncu_by_function.py report.ncu-rep source.cu [top_lines]
Needs a report captured with --import-source on and a -lineinfo build."""
import csv, io, re, subprocess, sys

rep, srcfile = sys.argv[1], sys.argv[2]
top_lines = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Line No"][0]
hdr = rows[hi]
first_file = [r[1] for r in rows[:hi] if r and r[0] == "File Path"]
curfile = first_file[0] if first_file else srcfile
isamp, iex = hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]


def num(x):
    try:
        return int(x)
    except ValueError:
        return 0


import os
lines = []
for r in rows[hi + 1:]:
    if not r or r[0] == "":
        continue
    if r[0] == "File Path":           # the report has one section per source file (headers of cooperative groups, sm_*_intrinsics ...)
        curfile = r[1]
        continue
    try:
        ln = int(r[0])
    except ValueError:
        continue
    own = os.path.basename(curfile) == os.path.basename(srcfile)
    lines.append((ln if own else -1, (r[1].strip() if own else os.path.basename(curfile) + ":" + r[0] + " " + r[1].strip()), num(r[isamp]), num(r[iex]),
                  {h: num(r[i]) for i, h in stall_cols}, own, os.path.basename(curfile)))
tot = sum(l[2] for l in lines) or 1
src = open(srcfile).read().split("\n")
func_at, cur = [], "?"
for l in src:
    m = re.match(r"^(?:__device__|__global__|static|__host__|template|inline|int |void |double ).*?(\w+)\s*\(", l)
    if m and not l.startswith(" ") and not l.startswith("template"):
        cur = m.group(1)
    func_at.append(cur)
agg = {}
for ln, txt, s, ex, st, own, fb in lines:
    f = (func_at[ln - 1] if 0 <= ln - 1 < len(func_at) else "?") if own else "[" + fb + "]"
    a = agg.setdefault(f, [0, 0, {}])
    a[0] += s; a[1] += ex
    for k, v in st.items():
        a[2][k] = a[2].get(k, 0) + v
print(f"total samples {tot}")
print("| function | samples | % | warp instructions | top stall reasons |\n|---|---|---|---|---|")
for f, (s, ex, st) in sorted(agg.items(), key=lambda x: -x[1][0])[:24]:
    top = sorted(st.items(), key=lambda x: -x[1])[:4]
    print(f"| {f} | {s} | {100 * s / tot:.1f} | {ex} | " + ", ".join(f"{k[6:]} {v}" for k, v in top if v) + " |")
print("\n| line | samples | % | source |\n|---|---|---|---|")
for ln, txt, s, ex, st, own, fb in sorted(lines, key=lambda x: -x[2])[:top_lines]:
    top = sorted(st.items(), key=lambda x: -x[1])[:2]
    print(f"| {ln} | {s} | {100 * s / tot:.1f} | `{txt[:110]}` ({', '.join(k[6:] for k, v in top if v)}) |")
