#!/bin/bash
# Last GPU call of the round: A/B of programmatic dependent launch on the resize chain, then ORB parity tests and bench.py with the
# library's final defaults (+ PDL if it won).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout 150 python tools/orb_variants.py --steps 20 --timeout 45 --out $O/r02f_orb_variants.jsonl > $O/r02f_orb_variants.log 2>&1
pdl=$(python - <<'P'
import json
rs=[json.loads(l) for l in open("gpurun_out/r02f_orb_variants.jsonl") if l.strip()]
for r in rs: print(r["config"], "|", r.get("ms_per_step"), r.get("bit_identical_to_round1"), r.get("per_kernel_ms"), (r.get("error") or "")[-200:], file=__import__("sys").stderr)
ok=[r for r in rs if r.get("bit_identical_to_round1") and "ms_per_step" in r]
base=[r for r in ok if r["env"].get("SE2GPU_ORB_PDL")=="0" and "SE2GPU_ORB_FAST_TMA" not in r["env"]]
pd=[r for r in ok if r["env"].get("SE2GPU_ORB_PDL")=="1"]
print(1 if base and pd and pd[0]["ms_per_step"] < base[0]["ms_per_step"]*0.997 else 0)
P
)
echo "PDL chosen: $pdl" | tee $O/r02f_pdl_choice.txt
export SE2GPU_ORB_PDL=$pdl
timeout 120 python -m pytest tests/test_orb_gpu.py tests/test_cpp_shim.py -m gpu -x -q > $O/r02f_pytest_orb.log 2>&1; echo "pytest rc=$?" | tee -a $O/r02f_pytest_orb.log
tail -2 $O/r02f_pytest_orb.log
timeout 200 python bench.py --no-c5 > $O/r02f_bench_n1.json 2> $O/r02f_bench_n1.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.load(open("gpurun_out/r02f_bench_n1.json"))
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["per_kernel_ms"])
P
