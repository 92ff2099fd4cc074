"""Measurement behind DESIGN.md's lane statistics: rollout lengths and the share of lane-steps that
belong to agents that still move, per workload (run on a GPU box)."""
import sys
import numpy as np
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import bench
for kind in ("M", "W"):
    w = bench.build_workload(kind, 0, 0, None)
    eng, leaves = w["engine"], w["leaves"]
    n = 64
    out = eng.rollout(leaves, n, seed=1000, max_steps=w["max_steps"], detail=True)
    steps = out["ro_steps"].astype(np.int64)
    tot_lane_steps = steps.sum() * eng.M
    print(kind, "rollouts", steps.size, "mean steps %.2f" % steps.mean(), "agent_steps", int(out["agent_steps"].sum()),
          "lane-steps", int(tot_lane_steps), "moving fraction %.3f" % (out["agent_steps"].sum() / tot_lane_steps),
          "hist", np.bincount(steps, minlength=21)[:21].tolist())
    fl = out["ro_flags"]
    print("  final valid fraction per agent", ((fl & 2) != 0).mean(axis=0).round(3).tolist())
