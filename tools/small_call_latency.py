#!/usr/bin/env python
"""Latency of small rollout calls, the shape a host tree issues per round (a handful of new
leaves x 16 rollouts): wall clock per C-ABI call on host buffers and device time of its kernels.
One JSON line."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import planner_rules_mcts_b200 as brm  # noqa: E402
from planner_rules_mcts_b200 import scenarios  # noqa: E402

res = {}
for name, scn in (("merge", scenarios.merge_scenario(seed=1000)), ("highway", scenarios.highway_scenario(seed=1000))):
    eng = brm.RulesMctsEngine([scn], device=0)
    for L, n in ((1, 1), (1, 16), (8, 16), (32, 16), (64, 64), (256, 64)):
        A = len(scn.agents)
        leaves = eng.make_batch(np.broadcast_to(scn.initial_state, (L, A, 4)).copy())
        for i in range(5):
            eng.rollout(leaves, n, seed=1000, rollout_base=i * L * n, max_steps=20)
        eng.engine.set_kernel_timing(True)
        eng.engine.kernel_time(reset=True)
        K = 50
        t0 = time.perf_counter()
        for i in range(K):
            eng.rollout(leaves, n, seed=1000, rollout_base=(10 + i) * L * n, max_steps=20)
        wall = (time.perf_counter() - t0) / K
        ms, _ = eng.engine.kernel_time(reset=True)
        eng.engine.set_kernel_timing(False)
        res["%s %dx%d" % (name, L, n)] = dict(wall_us=wall * 1e6, kernel_us=ms * 1e3 / K)
print(json.dumps(res))
