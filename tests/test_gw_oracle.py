"""Pins the GridWorld CPU oracle (oracle/gw_oracle.c) against
  (1) the 23 known answers of /root/reference/gridworld/label_evaluator/test/label_test.cpp,
  (2) golden traces generated from the reference's own code (tests/golden/gw_golden.npz),
  (3) the reference's own code run live (oracle/_ref, only where it was built)."""
import os

import numpy as np
import pytest

from oracle import gw as ogw
from tests.gw_common import random_states
from tests.golden.make_gw_golden import scenario

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "gw_golden.npz")
HAVE_REF = os.path.exists(os.path.join(ogw.HERE, "_ref", "libgw_ref.so"))

# label_test.cpp:12-64 -- EvaluatorLabelAtPosition("position", 5): (x, last_action, expected)
AT_POSITION = [(5, 1, True), (6, 2, True), (4, -2, True), (4, -1, False), (6, 1, False),
               (7, 1, False), (3, -1, False)]
# label_test.cpp:66-163 -- EvaluatorLabelCollision("collision", 0), both perspectives:
# (ego x,last,lane, other x,last,lane, expected)
COLLISION = [
    ((0, 1, 1), (0, 1, 2), True),   # :70-75 default agents: x=0, last=FORWARD, lanes 1 and 2, both on the merge point 0
    ((5, 2, 1), (4, -1, 1), True),
    ((5, 2, 1), (4, 1, 1), True),
    ((5, 2, 1), (4, 2, 1), True),
    ((5, 2, 1), (3, 1, 1), False),
    ((5, 2, 1), (5, 1, 0), False),
    ((1, 0, 1), (1, 0, 1), True),
    ((14, 2, 0), (13, 0, 0), True),
]


def impls():
    out = [ogw.GwOracle("port")]
    if HAVE_REF:
        out.append(ogw.GwOracle("ref"))
    return out


def test_label_known_answers(built):
    n = 0
    for orc in impls():
        for x, last, want in AT_POSITION:
            assert orc.at_position(x, last, 5) == want
            n += 1
        for e, o, want in COLLISION:
            assert orc.label_collision(0, *e, *o) == want
            assert orc.label_collision(0, *o, *e) == want
            n += 2
    assert n % 23 == 0  # 7 + 16 assertions, as label_test.cpp


def test_action_cost_probed_values(built):
    # WAIT after FORWARD = -6.5, FORWARD after WAIT = -1.5, FORWARD after FORWARD = 0
    port = ogw.GwOracle("port")
    cfg = ogw.base_test_env_config(terminal_depth=100)
    ag, term = ogw.base_test_env_state()
    o = port.replay(cfg, ag, term, 0, np.array([[1, 0, 0], [0, 0, 0], [0, 0, 0]], np.uint8))
    assert o["rewards"][0, 0, 2] == -6.5 and o["rewards"][1, 0, 2] == -1.5 and o["rewards"][2, 0, 2] == 0.0


def test_port_matches_golden_from_reference(built):
    g = np.load(GOLDEN)
    port = ogw.GwOracle("port")
    for name in ["base3_reset", "base3_stay", "merge4_reset"]:
        cfg, A = scenario(name)
        ag, term, acts = g[name + "/agents0"], g[name + "/term0"], g[name + "/actions"]
        n_viol = 0
        for b in range(ag.shape[0]):
            o = port.replay(cfg, ag[b], term[b], 0, acts[b])
            assert o["steps"] == g[name + "/steps"][b]
            for k in ["agents", "term", "q", "viol", "rewards", "labels", "term_reward"]:
                assert np.array_equal(o[k], g[name + "/" + k][b]), (name, b, k)
            n_viol += int(o["viol"].max())
        assert n_viol > 0, "golden set should exercise rule violations"


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref not built (needs /root/reference)")
def test_port_matches_reference_live(built):
    port, ref = ogw.GwOracle("port"), ogw.GwOracle("ref")
    rng = np.random.default_rng(3)
    for A, depth in [(2, 9), (3, 12), (4, 10), (5, 8)]:
        for reset in (True, False):
            cfg = ogw.base_test_env_config(A, terminal_depth=depth, reset_on_violation=reset)
            agents, term = random_states(rng, 150, A, p_terminal=0.15)
            for b in range(agents.shape[0]):
                acts = rng.integers(0, 3, size=(depth + 2, A)).astype(np.uint8)
                a = port.replay(cfg, agents[b], term[b], 0, acts)
                r = ref.replay(cfg, agents[b], term[b], 0, acts)
                for k in a:
                    assert np.array_equal(a[k], r[k]), (A, b, k)


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref not built (needs /root/reference)")
def test_rollouts_port_matches_reference_live(built):
    port, ref = ogw.GwOracle("port"), ogw.GwOracle("ref")
    cfg = ogw.base_test_env_config()
    ag, term = ogw.base_test_env_state()
    a = port.rollout(cfg, ag, term, 0, 1000, 0, 3000, 30, 0.95, n_threads=3)
    r = ref.rollout(cfg, ag, term, 0, 1000, 0, 3000, 30, 0.95, n_threads=2)
    assert a["agent_steps"] == r["agent_steps"] > 0
    assert np.array_equal(a["viol_sum"], r["viol_sum"])
    np.testing.assert_allclose(a["returns_sum"], r["returns_sum"], rtol=1e-12)


def test_terminal_agents_and_quirk_a1(built):
    """Appendix A-1: two agents colliding on the same non-zero lane before the merge point
    end on different lanes (only the first one is moved to lane 0), so `collision` is false."""
    port = ogw.GwOracle("port")
    cfg = ogw.base_test_env_config(terminal_depth=100)
    agents = np.array([[2, 1, 1, 1], [3, 1, 1, 1], [-20, 1, 3, 3]], np.int32)
    term = np.zeros(3, np.uint8)
    # agent 0 forward onto agent 1 who waits
    o = port.replay(cfg, agents, term, 0, np.array([[0, 1, 1]], np.uint8))
    assert o["agents"][0, 0, 0] == o["agents"][0, 1, 0] == 3
    assert o["agents"][0, 0, 2] == 0 and o["agents"][0, 1, 2] == 1
    assert (o["labels"][0, :2] & 1).tolist() == [0, 0]
    assert o["term"][0].tolist() == [0, 0, 0]
