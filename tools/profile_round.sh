#!/bin/bash
# Runs ON THE GPU BOX (under gpurun), one capture per call:
#   tools/profile_round.sh M|G|S|W    plain run first, then ONE `ncu --set full` launch of the rollout kernel
#   tools/profile_round.sh launches   plain run, then the launch list of a decision of M
# Outputs go to gpurun_out/ ; summarised afterwards with tools/make_inst_counts.py <tag>.
set -u
W=${1:-M}
mkdir -p gpurun_out
if [ "$W" = "launches" ]; then
  CMD="python bench.py --workload M --steps 4 --warmup 3 --no-cpu-baseline"
  timeout 300 $CMD > gpurun_out/plain_M_l.json 2> gpurun_out/plain_M_l.err && \
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_M.csv \
      $CMD > gpurun_out/ncu_launch_M.log 2>&1
  echo "launch list rc=$?"
  exit 0
fi
CMD="python bench.py --workload $W --steps 4 --warmup 3 --no-cpu-baseline"
K='regex:cw_pool_kernel<\(int\)0'; [ "$W" = "G" ] && K='regex:gw_rollout_kernel'
timeout 300 $CMD > gpurun_out/plain_$W.json 2> gpurun_out/plain_$W.err && \
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "$K" -s 4 -c 1 \
    -f -o gpurun_out/prof_$W $CMD > gpurun_out/ncu_full_$W.log 2>&1
echo "$W full rc=$?"
tail -3 gpurun_out/ncu_full_$W.log
