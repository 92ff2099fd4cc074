#!/usr/bin/env python
"""Key metrics of one ncu report as JSON: ncu_summary.py <report.ncu-rep> [kernel-substring]"""
import csv
import io
import json
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__warps_active.avg.per_cycle_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.max",
        "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shared_st.sum",
        "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
        "sm__inst_executed_pipe_fp64.sum", "smsp__cycles_active.avg"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
res = []
for vals in rows[2:]:
    d = {}
    for i, h in enumerate(hdr):
        if h in ("Kernel Name",) or h in WANT:
            d[h] = vals[i] + ((" " + units[i]) if units[i] else "")
    if len(sys.argv) < 3 or sys.argv[2] in d.get("Kernel Name", ""):
        res.append(d)
print(json.dumps(res, indent=1))
