#!/bin/bash
# One GPU call: TMA probes, A/B of the ORB kernel variants (bit-identity + time), then the full GPU test suite, bench.py and ncu
# captures with the winning settings. Every step has its own timeout; everything lands in gpurun_out/.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
{
for cfg in "0 128 81 16 7 2" "1 128 81 16 7 2" "3 128 81 32 9 1" "2 128 81 48 3 3" "1 112 68 624 400 1" "0 128 81 8 7 2"; do
  echo -n "cfg [$cfg]: "; timeout 30 tools/tma_probe $cfg | tail -1
done
} > $O/r02c_tma_probe.log 2>&1
cat $O/r02c_tma_probe.log
timeout 300 python tools/orb_variants.py --steps 20 --out $O/r02c_orb_variants.jsonl > $O/r02c_orb_variants.log 2>&1
python tools/choose_orb_env.py $O/r02c_orb_variants.jsonl > $O/r02c_chosen_env.sh 2> $O/r02c_choose.err
cat $O/r02c_chosen_env.sh
source $O/r02c_chosen_env.sh
timeout 300 python -m pytest tests -m gpu -x -q > $O/r02c_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/r02c_pytest_gpu.log
tail -3 $O/r02c_pytest_gpu.log
timeout 420 python bench.py > $O/r02c_bench_n1.json 2> $O/r02c_bench_n1.err; echo "bench rc=$?"
timeout 200 ncu --set full --import-source on --clock-control none -k regex:'orb_fast_cells|orb_orient|orb_resize|orb_pyr0|orb_blur|orb_select' --launch-skip 13 -c 13 -f -o $O/r02c_prof_orb python tools/prof_targets.py orb > $O/r02c_ncu_orb.log 2>&1; echo "ncu rc=$?"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02c_launches_bench_quick.csv python bench.py --steps 2 --warmup 1 --quick --no-c5 > $O/r02c_bench_under_ncu.log 2>&1; echo "launch list rc=$?"
ls -la $O | tail -12
