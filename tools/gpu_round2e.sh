#!/bin/bash
# GPU call: A/B of the blur schedule and of the shared-memory-lean resize (bit-identity + time), ORB parity tests with the resize variant.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout 360 python tools/orb_variants.py --steps 20 --out $O/r02e_orb_variants.jsonl > $O/r02e_orb_variants.log 2>&1
python - <<'P'
import json
for l in open("gpurun_out/r02e_orb_variants.jsonl"):
    r=json.loads(l); print(r["config"], "|", r.get("ms_per_step"), r.get("bit_identical_to_round1"), r.get("per_kernel_ms"), (r.get("error") or "")[-200:])
P
SE2GPU_ORB_RESIZE_W=1 SE2GPU_ORB_BLUR_B_AFTER_FAST=1 timeout 200 python -m pytest tests/test_orb_gpu.py tests/test_cpp_shim.py -m gpu -x -q > $O/r02e_pytest_orb.log 2>&1; echo "pytest rc=$?" | tee -a $O/r02e_pytest_orb.log
tail -3 $O/r02e_pytest_orb.log
