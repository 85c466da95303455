#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
{
for cfg in "0 16 16" "0 32 32" "0 64 64" "0 64 65" "0 64 81" "0 64 128" "0 64 256" "0 128 16" "0 128 31" "0 128 32" "0 128 38" "0 128 64" "0 80 38" "0 96 38" "0 112 38" "0 256 8" "0 144 8" "0 16 256" "0 128 81 0 0 0" "0 128 81 16 16 0" "1 64 38" "2 64 38" "3 64 38" "1 128 16" ; do
  echo -n "cfg [$cfg]: "; timeout 30 tools/tma_probe $cfg | tail -1
done
} > $O/r02b_tma_probe.log 2>&1
cat $O/r02b_tma_probe.log
