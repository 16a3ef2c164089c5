#!/bin/bash
# One `ncu --set full` capture per driver kernel on the small single-launch cases of tools/prof_case.py (run on the GPU box).
# Reports land in gpurun_out/; tools/ncu_traffic_summary.py turns them into profiles/r02_traffic.json.
set -u
mkdir -p gpurun_out
for c in c1 c2 c3 c4 c5; do
  python tools/prof_case.py $c > gpurun_out/prof_$c.log 2>&1 || { echo "prof_case $c failed"; continue; }
  k=$(python -c "from unbiasedmcmc_b200.workloads import KERNEL_OF; print(KERNEL_OF['$c'])")
  ncu --set full --import-source on --clock-control none -k regex:$k -c 1 -f -o gpurun_out/r02_full_$c python tools/prof_case.py $c > gpurun_out/ncu_$c.log 2>&1
  tail -1 gpurun_out/ncu_$c.log
done
