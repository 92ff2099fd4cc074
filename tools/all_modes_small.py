#!/usr/bin/env python
"""A small pass over every kernel mode (rollout in both launch shapes, execute, replay, rules,
queries, GridWorld), sized for a debugger or a sanitizer run."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import planner_rules_mcts_b200 as brm  # noqa: E402
from planner_rules_mcts_b200 import scenarios  # noqa: E402

scn = scenarios.merge_scenario(seed=1000)
eng = brm.RulesMctsEngine([scn], device=0)
root = eng.initial_batch()
out = eng.rollout(root, 700, seed=3, max_steps=20, detail=True)       # large shape, pools refill
print("merge rollouts", int(out["agent_steps"].sum()))
acts = np.zeros((eng.M, 1), np.uint8)
tr = eng.replay(root, np.zeros((1, 5, eng.M), np.uint8))                  # REPLAY
print("replay steps", tr["steps_out"] if isinstance(tr, dict) else "ok")
nxt, rew = eng.execute(root, acts)                                      # EXECUTE
print("execute", rew.ravel()[:3])
print("queries", eng.get_num_actions(nxt).ravel(), eng.terminal_reward(nxt).ravel()[:2])
print("rules", eng.evaluate_rules(nxt)[1].ravel()[:2])
scns = [scenarios.intersection_scenario(seed=1000, scenario_id=k) for k in range(3)]
e2 = brm.RulesMctsEngine(scns, device=0)
o2 = e2.rollout(e2.initial_batch(), 150, seed=5, detail=True)           # small shape, shared leaves
print("intersection rollouts", int(o2["agent_steps"].sum()))
g = brm.GridWorldEngine(device=0)
o3 = g.rollout(g.base_test_env_batch(2), 64, seed=1000, max_steps=30, discount_factor=0.95)
print("gridworld", o3["returns_sum"].ravel()[:3])
