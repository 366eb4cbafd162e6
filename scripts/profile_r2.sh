#!/bin/bash
# ncu evidence of round 2 (1 GPU): launch list of a short bench run + one --set full capture per hot kernel
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv $B > gpurun_out/r02_launches_bench.log 2>&1
for k in k_update_steps5 k_assign_tc3 k_apply_tc k_stats_tc; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -f -o gpurun_out/r02_$k $B > gpurun_out/r02_ncu_$k.log 2>&1
  tail -2 gpurun_out/r02_ncu_$k.log
done
ls -la gpurun_out/*.ncu-rep
