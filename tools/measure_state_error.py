"""Measurement behind DESIGN.md's error figures: largest position / reward difference between
the CUDA replay and the FP64 oracle on random action sequences (run on a GPU box)."""
import sys
import numpy as np
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import planner_rules_mcts_b200 as brm
from planner_rules_mcts_b200 import scenarios
from oracle import cw as ocw
from tests.test_cw_gpu import replay_both
orc = ocw.CwOracle()
for name, scn in [("highway", scenarios.highway_scenario()), ("test_world", scenarios.straight_road_test_world(
        n_others=2, multi_agent=True, primitives=[(0.0, 0.0), (3.0, 0.05), (-2.0, -0.05), (4.0, 0.0)]))]:
    eng = brm.RulesMctsEngine([scn], device=0)
    rng = np.random.default_rng(11)
    B, T = 256, 22
    n_act = len(scn.agents[0].primitives)
    acts = rng.integers(0, n_act, size=(B, T, scn.n_mcts_agents)).astype(np.uint8)
    tr, outs, _ = replay_both(eng, orc, scn, acts)
    e_pos = e_rew = 0.0
    for b, o in enumerate(outs):
        k = o["steps"]
        if k == 0: continue
        d = np.abs(tr["agents_out"][b, :k].astype(np.float64) - o["states"][:k])
        e_pos = max(e_pos, d[..., :2].max())
        e_rew = max(e_rew, np.abs(tr["rewards_out"][b, :k].astype(np.float64) - o["rewards"][:k]).max())
    print(name, "max |pos err| %.3g m, max |reward err| %.3g" % (e_pos, e_rew))
