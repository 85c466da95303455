#!/bin/bash
# Final check of the committed defaults: the whole GPU test suite and smoke().
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
timeout 150 python -m pytest tests -m gpu -x -q > $O/r02g_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/r02g_pytest_gpu.log
tail -3 $O/r02g_pytest_gpu.log
timeout 40 python __graft_entry__.py --smoke > $O/r02g_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/r02g_smoke.log
