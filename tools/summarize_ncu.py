#!/usr/bin/env python
"""Summarise an ncu report (ncu --set full) into a small markdown table for profiles/."""
import csv
import io
import subprocess
import sys

WANT = [("gpu__time_duration.sum", "duration"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__registers_per_thread", "regs"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_%"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ_%"),
        ("smsp__inst_executed.sum", "warp_inst"), ("smsp__thread_inst_executed_per_inst_executed.ratio", "lanes/inst"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex_%"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_%")]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    lines = [f"# ncu summary of `{rep}`", "", "| kernel | " + " | ".join(n for _, n in WANT) + " |", "|---|" + "---|" * len(WANT)]
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0].replace("<unnamed>::", "")
        vals = []
        for key, _ in WANT:
            if key in idx and r[idx[key]] != "":
                v, u = r[idx[key]], units[idx[key]]
                try:
                    vals.append(f"{float(v.replace(',', '')):.4g} {u}".strip())
                except ValueError:
                    vals.append(v)
            else:
                vals.append("-")
        lines.append(f"| {name} | " + " | ".join(vals) + " |")
    open(out, "w").write("\n".join(lines) + "\n")
    print("wrote", out, len(rows) - 2, "launches")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
