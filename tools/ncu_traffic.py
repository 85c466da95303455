#!/usr/bin/env python
"""Builds profiles/rNN_ncu_traffic.json (what bench.py copies into roofline.traffic / roofline.issue) from ncu reports.
usage: ncu_traffic.py out.json report1.ncu-rep [report2.ncu-rep ...]
Per kernel name (first launch of each name; orb_resize launches are suffixed _level1.. in order; orb_blur launches are summed
as one logical launch group): dram__bytes_read.sum + dram__bytes_write.sum, issue-slot utilisation, lanes per instruction,
warp instructions."""
import csv, io, json, subprocess, sys

out, reps = sys.argv[1], sys.argv[2:]
res = {"_comment": "dram__bytes_read.sum + dram__bytes_write.sum per launch from `ncu --set full --clock-control none` captures of round 2 (" + ", ".join(reps) + "); bench.py copies these into roofline.traffic / roofline.issue"}
for rep in reps:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}

    def val(r, key):
        v = r[ix[key]].replace(",", "")
        u = units[ix[key]]
        f = float(v)
        mult = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}.get(u, 1.0)
        return f * mult
    nres, nblur = 0, 0
    for r in rows[2:]:
        name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "").strip()
        name = name.split("::")[-1].split("<")[0]            # drop namespaces ("<unnamed>::", "se2band::<unnamed>::") and template arguments
        traffic = val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")
        issue = {"issue_slot_pct": float(r[ix["smsp__issue_active.avg.pct_of_peak_sustained_active"]]),
                 "lanes_per_inst": float(r[ix["smsp__thread_inst_executed_per_inst_executed.ratio"]]),
                 "warp_inst_per_launch": float(r[ix["smsp__inst_executed.sum"]].replace(",", "")), "source": rep,
                 "duration_us": val(r, "gpu__time_duration.sum") / (1e3 if units[ix["gpu__time_duration.sum"]] in ("ns", "nsecond") else 1.0)}
        if name == "orb_resize":
            nres += 1
            name = f"orb_resize_level{nres}"
        if name == "orb_blur":
            nblur += 1
            if "orb_blur" in res:
                res["orb_blur"] += traffic
                res["issue:orb_blur"]["warp_inst_per_launch"] += issue["warp_inst_per_launch"]
                continue
        if name in res:
            continue
        res[name] = traffic
        res["issue:" + name] = issue
band = [k for k in ("band_part_factor", "band_sep_solve", "band_part_back") if k in res]
if band:   # the C5 leg's reduced solve is these three launches; bench.py looks the group up as c5:ba_chol_solve
    res["c5:ba_chol_solve"] = sum(res[k] for k in band)
if "orb_resize_level1" in res:
    res["pyramid"] = res.get("orb_pyr0", 0) + sum(v for k, v in res.items() if k.startswith("orb_resize_level"))
    res["issue:pyramid"] = res["issue:orb_resize_level1"]
json.dump(res, open(out, "w"), indent=1)
print("wrote", out, [k for k in res if not k.startswith("issue:") and k != "_comment"])
