#!/bin/bash
# One `ncu --set full` capture per driver kernel on the small single-launch cases of tools/prof_case.py (run on the GPU box).
# The reports stay in /tmp on the box (five of them exceed what gpurun copies back); their raw pages (and the source pages of
# the kernels rewritten this round) come back as CSV in gpurun_out/; tools/ncu_traffic_summary.py turns them into
# profiles/r02_traffic.json and profiles/r02_kernel_metrics.md.
set -u
mkdir -p gpurun_out
for c in c1 c2 c3 c4 c5; do
  python tools/prof_case.py $c > gpurun_out/prof_$c.log 2>&1 || { echo "prof_case $c failed"; continue; }
  k=$(python -c "from unbiasedmcmc_b200.workloads import KERNEL_OF; print(KERNEL_OF['$c'])")
  ncu --set full --import-source on --clock-control none -k regex:$k -c 1 -f -o /tmp/r02_full_$c python tools/prof_case.py $c > gpurun_out/ncu_$c.log 2>&1
  tail -1 gpurun_out/ncu_$c.log
  ncu -i /tmp/r02_full_$c.ncu-rep --page raw --csv > gpurun_out/r02_full_${c}_raw.csv 2>/dev/null
  if [ $c = c3 ] || [ $c = c4 ]; then
    ncu -i /tmp/r02_full_$c.ncu-rep --page source --print-source cuda,sass --csv 2>/dev/null | gzip > gpurun_out/r02_full_${c}_source.csv.gz
  fi
done
ls -la gpurun_out | tail -12
