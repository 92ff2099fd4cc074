#!/bin/bash
# Runs ON THE GPU BOX (under gpurun): plain run first, then ncu on the identical command.
# Outputs go to gpurun_out/ ; summarised afterwards with tools/make_inst_counts.py.
set -u
for W in M G S W; do
  CMD="python bench.py --workload $W --steps 4 --warmup 3 --no-cpu-baseline"
  K=cw_rollout_kernel; [ "$W" = "G" ] && K=gw_rollout_kernel
  timeout 300 $CMD > gpurun_out/plain_$W.json 2> gpurun_out/plain_$W.err && \
  ncu --set full --clock-control none --import-source on -k regex:$K -s 4 -c 1 -o gpurun_out/prof_$W \
      $CMD > gpurun_out/ncu_full_$W.log 2>&1
  echo "$W full rc=$?"
done
CMD="python bench.py --workload M --steps 4 --warmup 3 --no-cpu-baseline"
timeout 300 $CMD > gpurun_out/plain_M_l.json 2> gpurun_out/plain_M_l.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_M.csv \
    $CMD > gpurun_out/ncu_launch_M.log 2>&1
echo "launch list rc=$?"
