"""Reads gpurun_out/orb_variants.jsonl (tools/orb_variants.py) and prints `export VAR=value` lines for the fastest setting of every
ORB kernel knob whose output was bit-identical to the round-1 kernels; knobs without a valid faster variant keep the round-1 value."""
import json
import sys

recs = [json.loads(l) for l in open(sys.argv[1]) if l.strip()]
ok = [r for r in recs if r.get("bit_identical_to_round1") and "per_kernel_ms" in r]
base = ok[0] if ok and ok[0]["config"] == "round-1 kernels" else None
env = {"SE2GPU_ORB_FAST_TMA": "0", "SE2GPU_ORB_ORIENT_BATCH": "0"}
if base:
    def single(r, keys):   # configurations that differ from the baseline only in `keys`
        return all(r["env"][k] == base["env"][k] for k in env if k not in keys)
    best = base["per_kernel_ms"]["orb_fast_cells"]
    for r in ok:
        if single(r, ("SE2GPU_ORB_FAST_TMA",)) and r["per_kernel_ms"]["orb_fast_cells"] < best * 0.995:
            best = r["per_kernel_ms"]["orb_fast_cells"]; env["SE2GPU_ORB_FAST_TMA"] = r["env"]["SE2GPU_ORB_FAST_TMA"]
    best = base["per_kernel_ms"]["orb_orient_describe"]
    for r in ok:
        if single(r, ("SE2GPU_ORB_ORIENT_BATCH",)) and r["per_kernel_ms"]["orb_orient_describe"] < best * 0.995:
            best = r["per_kernel_ms"]["orb_orient_describe"]; env["SE2GPU_ORB_ORIENT_BATCH"] = r["env"]["SE2GPU_ORB_ORIENT_BATCH"]
for k, v in env.items():
    print(f"export {k}={v}")
