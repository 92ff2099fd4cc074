#!/usr/bin/env python
"""Builds profiles/inst_counts.json (read by bench.py for the issue roofline) and the
per-workload ncu summaries from gpurun_out/prof_<W>.ncu-rep + gpurun_out/plain_<W>.json
(the plain, un-profiled run of the same command line: source of agent-steps per launch)."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TAG = sys.argv[1] if len(sys.argv) > 1 else "r01"
out = {}
for w in "MGSW":
    rep = os.path.join(ROOT, "gpurun_out", "prof_%s.ncu-rep" % w)
    plain = os.path.join(ROOT, "gpurun_out", "plain_%s.json" % w)
    if not (os.path.exists(rep) and os.path.exists(plain)):
        continue
    line = [l for l in open(plain) if l.startswith("{")][-1]
    b = json.loads(line)
    steps = b["config"]["agent_steps_per_decision_per_gpu"]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, vals = rows[0], rows[2]
    get = lambda k: float(vals[hdr.index(k)].replace(",", ""))
    unit = lambda k: rows[1][hdr.index(k)]
    winst = get("smsp__inst_executed.sum")
    tinst = winst * get("smsp__thread_inst_executed_per_inst_executed.ratio")
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    dram = get("dram__bytes_read.sum") * scale[unit("dram__bytes_read.sum")] + \
        get("dram__bytes_write.sum") * scale[unit("dram__bytes_write.sum")]
    dur_unit = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[unit("gpu__time_duration.sum")]
    out[w] = dict(
        warp_inst_per_agent_step=winst / steps, thread_inst_per_agent_step=tinst / steps,
        warp_inst_per_launch=winst, thread_inst_per_launch=tinst, agent_steps_per_launch=steps,
        avg_active_lanes=tinst / winst, dram_bytes_per_launch=dram,
        ncu_duration_ms=get("gpu__time_duration.sum") * dur_unit,
        issue_active_pct=get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        warps_active_per_scheduler=get("smsp__warps_active.avg.per_cycle_active"),
        warps_eligible_per_scheduler=get("smsp__warps_eligible.avg.per_cycle_active"),
        registers_per_thread=get("launch__registers_per_thread"), grid_size=get("launch__grid_size"),
        block_size=get("launch__block_size"), waves_per_sm=get("launch__waves_per_multiprocessor"),
        source="ncu --set full --clock-control none, one launch of `python bench.py --workload %s --steps 4 "
               "--warmup 3 --no-cpu-baseline` (profiles/%s_ncu_full_rollout_%s.json); agent-steps per launch = "
               "mean of the plain run of the same command" % (w, TAG, w))
    summ = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep],
                          capture_output=True, text=True).stdout
    open(os.path.join(ROOT, "profiles", "%s_ncu_full_rollout_%s.json" % (TAG, w)), "w").write(summ)
sys.path.insert(0, ROOT)
import bench  # noqa: E402
out["csrc_sha16"] = bench.kernel_source_hash()   # bench.py flags the counts as stale when the kernels change
json.dump(out, open(os.path.join(ROOT, "profiles", "inst_counts.json"), "w"), indent=1)
for w, d in out.items():
    if not isinstance(d, dict):
        continue
    print(w, "warp-inst/agent-step %.1f  thread-inst/agent-step %.0f  lanes %.1f  issue %.1f%%  dur %.3f ms  dram %.0f B" % (
        d["warp_inst_per_agent_step"], d["thread_inst_per_agent_step"], d["avg_active_lanes"], d["issue_active_pct"],
        d["ncu_duration_ms"], d["dram_bytes_per_launch"]))
