#!/usr/bin/env python
"""Shape sensitivity of the multi-scenario rollout path (config W: one leaf per intersection
scenario, CTA-wide scenario blobs): the same 32768 rollouts per launch as S scenarios x n rollouts
each.  Few large leaves show the kernel without the per-leaf drain, many small ones the cost of
moving from scenario to scenario.  Device-timed with the engine's CUDA events; one JSON line."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import planner_rules_mcts_b200 as brm  # noqa: E402
from planner_rules_mcts_b200 import scenarios  # noqa: E402

TOTAL = 32768
res = {}
for S in (512, 128, 32, 8):
    n = TOTAL // S
    scns = [scenarios.intersection_scenario(seed=1000, scenario_id=k) for k in range(S)]
    eng = brm.RulesMctsEngine(scns, device=0)
    d = eng.initial_batch().cuda(0)
    out = dict(returns_sum=torch.zeros((S, eng.M, eng.V), dtype=torch.float64, device="cuda"),
               viol_sum=torch.zeros((S, eng.M, eng.R), dtype=torch.int64, device="cuda"),
               agent_steps=torch.zeros(S, dtype=torch.int64, device="cuda"))
    for i in range(3):
        eng.rollout(d, n, seed=1000, rollout_base=i * TOTAL, max_steps=20, discount_factor=0.9, out=out)
    eng.engine.synchronize()
    eng.engine.set_kernel_timing(True)
    eng.engine.kernel_time(reset=True)
    steps, K = 0, 10
    for i in range(K):
        eng.rollout(d, n, seed=1000, rollout_base=(10 + i) * TOTAL, max_steps=20, discount_factor=0.9, out=out)
        eng.engine.synchronize()
        steps += int(out["agent_steps"].sum().item())
    ms, _ = eng.engine.kernel_time(reset=True)
    eng.engine.set_kernel_timing(False)
    res["scenarios=%d" % S] = dict(scenarios=S, rollouts_per_scenario=n, kernel_ms=ms / K,
                                   agent_steps_per_s=steps / (ms * 1e-3))
    del eng
print(json.dumps(dict(workload="intersection sweep, 32768 rollouts per launch, 10 agents, 21 rule states", results=res)))
