#!/usr/bin/env python
"""Leaf-shape sensitivity of the fused rollout kernel on the merge workload: the same 262144
rollouts per decision as L leaves x n rollouts each, n from 1024 (the bench's shape) down to 1 (what
mvmcts::RandomHeuristic does: ONE rollout per expanded node).  Device-timed with the
engine's CUDA events, leaves resident on the GPU.  Prints one JSON line."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import planner_rules_mcts_b200 as brm  # noqa: E402
from planner_rules_mcts_b200 import scenarios  # noqa: E402

TOTAL = 262144
scn = scenarios.merge_scenario(seed=1000)
eng = brm.RulesMctsEngine([scn], device=0)
rng = np.random.default_rng(0)
res = {}
for n in (1024, 64, 16, 4, 1):
    L = TOTAL // n
    # distinct leaves: children of the root under random first actions, a little jitter on top
    states = np.broadcast_to(scn.initial_state, (L, 4, 4)).copy()
    states[:, :, 0] += rng.uniform(-0.5, 0.5, size=(L, 4))
    root = eng.make_batch(states)
    first = rng.integers(0, 3, size=(L, eng.M)).astype(np.uint8)
    leaves, _ = eng.execute(root, first.T.copy())
    d = leaves.cuda(0)
    out = dict(returns_sum=torch.zeros((L, 4, 3), dtype=torch.float64, device="cuda"),
               viol_sum=torch.zeros((L, 4, eng.R), dtype=torch.int64, device="cuda"),
               agent_steps=torch.zeros(L, dtype=torch.int64, device="cuda"))
    for i in range(3):
        eng.rollout(d, n, seed=1000, rollout_base=i * TOTAL, max_steps=20, out=out)
    eng.engine.synchronize()
    eng.engine.set_kernel_timing(True)
    eng.engine.kernel_time(reset=True)
    steps = 0
    K = 10
    for i in range(K):
        eng.rollout(d, n, seed=1000, rollout_base=(10 + i) * TOTAL, max_steps=20, out=out)
        eng.engine.synchronize()
        steps += int(out["agent_steps"].sum().item())
    ms, launches = eng.engine.kernel_time(reset=True)
    eng.engine.set_kernel_timing(False)
    res["n=%d" % n] = dict(leaves=L, rollouts_per_leaf=n, kernel_ms=ms / K, agent_steps_per_s=steps / (ms * 1e-3))
base = res["n=1024"]["agent_steps_per_s"]
for k, v in res.items():
    v["of_n1024"] = v["agent_steps_per_s"] / base
print(json.dumps(dict(workload="merge, 262144 rollouts per decision", results=res)))
