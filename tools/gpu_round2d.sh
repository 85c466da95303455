#!/bin/bash
# Final GPU call of the round: streaming host path chunk sweep, then the full GPU test suite, smoke() and bench.py with the library's
# (new) defaults and the best SE2GPU_ORB_SUBMIT_CHUNKS.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python tools/orb_e2e_submit.py $O/r02d_orb_e2e_submit.jsonl > $O/r02d_orb_e2e_submit.log 2>&1
cat $O/r02d_orb_e2e_submit.log
best=$(python - <<'P'
import json
rs=[json.loads(l) for l in open("gpurun_out/r02d_orb_e2e_submit.jsonl") if l.strip()]
ok=[r for r in rs if "ms_per_batch" in r]
ref=[r for r in ok if r["submit_chunks"]=="4"]
if ref: ok=[r for r in ok if r["sha256"]==ref[0]["sha256"]]
print(min(ok,key=lambda r:r["ms_per_batch"])["submit_chunks"] if ok else 4)
P
)
echo "best submit chunks: $best" | tee $O/r02d_best_submit_chunks.txt
export SE2GPU_ORB_SUBMIT_CHUNKS=$best
timeout 300 python -m pytest tests -m gpu -x -q > $O/r02d_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/r02d_pytest_gpu.log
tail -3 $O/r02d_pytest_gpu.log
timeout 120 python __graft_entry__.py --smoke > $O/r02d_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/r02d_smoke.log
timeout 420 python bench.py > $O/r02d_bench_n1.json 2> $O/r02d_bench_n1.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.load(open("gpurun_out/r02d_bench_n1.json"))
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "sync", d["e2e"]["sync"]["ms_per_step"])
P
