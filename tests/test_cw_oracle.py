"""Continuous-world CPU oracle (oracle/cw_oracle.cpp) against the only numeric pins the
reference holds for the BARK-backed state: test/single_agent_test.cpp:76-174 (rewards 0 /
-800 / -1000, terminal flags, goal reached by repeating action 0) and
test/multi_agent_test.cpp:51-140 (3 -> 1 actions after a collision, frozen positions).
Everything else about this oracle is "parity unpinned" (see oracle/cw_oracle.h)."""
import numpy as np
import pytest

from oracle import cw as ocw
from planner_rules_mcts_b200 import scenarios
from planner_rules_mcts_b200.ltl import RuleMonitor
from tests.cw_common import oracle_config

FL = lambda n: np.full(n, ocw.PRESENT | ocw.VALID, np.uint8)


@pytest.fixture(scope="module")
def orc(built):
    return ocw.CwOracle()


def test_single_agent_execute_pins(orc):
    scn = scenarios.straight_road_test_world()  # 1 ego + 1 other, rel_distance 7, v 5
    cfg = oracle_config(scn)
    s0, fl = scn.initial_state, FL(2)
    assert not orc.check_terminal(cfg, s0, fl, 20)
    # :136-143 action 0 (a=0, delta=0): not terminal, reward[0] == 0 incl. terminal reward
    o = orc.replay(cfg, s0, fl, 20, np.array([[0]], np.uint8))
    assert o["steps"] == 1 and not o["terminal"][0]
    assert abs(o["rewards"][0, 0, 0] + o["term_reward"][0, 0]) < 1e-5
    # :146-152 then the crazy action (50, 1): leaves the corridor -> terminal, -800
    o = orc.replay(cfg, s0, fl, 20, np.array([[0], [1]], np.uint8))
    assert o["steps"] == 2 and o["terminal"][1]
    total = o["rewards"][:2, 0, 0].sum() + o["term_reward"][0, 0]
    assert abs(total + 800.0) < 1e-5
    # :154-159 from the initial state, action 2 rear-ends the other agent -> terminal, -1000
    o = orc.replay(cfg, s0, fl, 20, np.array([[2]], np.uint8))
    assert o["steps"] == 1 and o["terminal"][0]
    assert abs(o["rewards"][0, 0, 0] + o["term_reward"][0, 0] + 1000.0) < 1e-5
    assert o["viol"][0, 0, 0] == 1 and not (o["flags"][0, 0] & ocw.VALID)
    # :161-173 repeating action 0 reaches the goal -> terminal with goal_reached
    o = orc.replay(cfg, s0, fl, 10000, np.zeros((200, 1), np.uint8))
    k = o["steps"]
    assert o["terminal"][k - 1] and (o["labels"][k - 1, 0, 0] >> 2) & 1


def test_multi_agent_collision_pins(orc):
    scn = scenarios.straight_road_test_world(n_others=2, ego_velocity=0.0, velocity_difference=-3.0,
                                             multi_agent=True, primitives=[(0.0, 0.0), (50.0, 1.0), (4.0, 0.0)],
                                             rules=[])
    scn.initial_state[1] = (10.0, -1.75, 0.0, 3.0)
    scn.initial_state[2] = (14.1, -1.75, 0.0, 3.0)   # test/multi_agent_test.cpp:100-101
    cfg = oracle_config(scn)
    fl = FL(3)
    # :126 Execute({0, 2, 0}) makes agents 2 and 3 collide
    o = orc.replay(cfg, scn.initial_state, fl, 20, np.array([[0, 2, 0], [0, 0, 0]], np.uint8))
    assert o["steps"] == 2
    f1 = o["flags"][0]
    assert (f1[0] & ocw.VALID) and not (f1[1] & ocw.VALID) and not (f1[2] & ocw.VALID)  # :128-133 -> 1 action
    # :136-139 collided agents stay where they are, the ego (v = 0) too
    assert np.array_equal(o["states"][1, 1], o["states"][0, 1])
    assert np.array_equal(o["states"][1, 2], o["states"][0, 2])


def test_reward_terms_follow_reference_formulas(orc):
    # GetActionCost (:202-248) and PotentialReward (:250-264) with non-default weights
    scn = scenarios.straight_road_test_world(n_others=0, primitives=[(2.0, 0.0), (0.0, 0.1)], rules=[])
    p = scn.params
    p.ACCELERATION_WEIGHT, p.RADIAL_ACCELERATION_WEIGHT = -1.5, -2.0
    p.DESIRED_VELOCITY_WEIGHT, p.LANE_CENTER_WEIGHT, p.POTENTIAL_WEIGHT = -5.0, -3.0, 32.0
    cfg = oracle_config(scn)
    for act in (0, 1):
        o = orc.replay(cfg, scn.initial_state, FL(1), 20, np.array([[act]], np.uint8))
        x0, y0, th0, v0 = scn.initial_state[0].astype(np.float64)
        x1, y1, th1, v1 = o["states"][0, 0]
        dt = p.PREDICTION_TIME_SPAN
        a = (v1 - v0) / dt
        cost = p.ACCELERATION_WEIGHT * a * a * dt
        a_lat = (th1 - th0) / dt * (0.5 * a * dt + v0)
        cost += p.RADIAL_ACCELERATION_WEIGHT * a_lat * a_lat * dt
        cost += p.DESIRED_VELOCITY_WEIGHT * abs(v0 - p.DESIRED_VELOCITY) * dt
        cost += p.LANE_CENTER_WEIGHT * abs(y1 + 1.75) * dt
        pot = lambda v: -p.POTENTIAL_WEIGHT * dt * abs(v - p.DESIRED_VELOCITY)
        want = cost + p.DISCOUNT_FACTOR * pot(v1) - pot(v0)
        assert abs(o["rewards"][0, 0, 0] - want) < 1e-9 * max(1.0, abs(want))
        assert abs(o["frenet"][0, 0, 2] - (y1 + 1.75)) < 1e-9 and o["frenet"][0, 0, 0] == 0


def test_euler_substeps(orc):
    # 15 explicit Euler sub-steps of 0.02 s (bark BehaviorMPContinuousActions, un-vendored)
    scn = scenarios.straight_road_test_world(n_others=0, primitives=[(2.0, 0.05)], rules=[])
    cfg = oracle_config(scn)
    o = orc.replay(cfg, scn.initial_state, FL(1), 20, np.array([[0]], np.uint8))
    x, y, th, v = scn.initial_state[0].astype(np.float64)
    k = np.tan(0.05) / 2.7
    h = 0.02
    for n in range(15):
        hh = (0.3 - 14 * 0.02) if n == 14 else h
        x, y, th, v = x + hh * v * np.cos(th), y + hh * v * np.sin(th), th + hh * v * k, v + hh * 2.0
    np.testing.assert_allclose(o["states"][0, 0], [x, y, th, v], rtol=1e-12)


def test_rule_slots_and_zip_on_merge(orc):
    scn = scenarios.merge_scenario(seed=1000)
    cfg = oracle_config(scn)
    assert orc.rules_per_agent(cfg) == 1 + 3
    r = orc.rollout(cfg, scn.initial_state, FL(4), 20, 1000, 0, 1500, 20, 0.9, n_threads=2)
    assert 0 < r["agent_steps"] <= 1500 * 20 * 4
    assert r["viol_sum"][:, 0].sum() > 0          # safe-distance violations happen
    assert r["viol_sum"][:, 1:].sum() > 0         # and so do zipper violations
    assert np.all(r["returns_sum"][:, 2] < 0)     # comfort / desired velocity costs
    w = scenarios.intersection_scenario()
    assert orc.rules_per_agent(oracle_config(w)) == 3 + 2 * 9


def test_agents_leaving_the_map_are_removed(orc):
    """bark "World::remove_agents_out_of_map": Predict removes an agent whose shape left the road
    polygon; an MCTS agent that vanished earns OUT_OF_MAP_WEIGHT at index 0 that step and nothing
    else (src/rules_mcts_state.cpp:121-125), has one action from then on (:129-141), earns no
    terminal reward (:154-156); without the ego the state is terminal (:267-270)."""
    prim = [(0.0, 0.0), (0.0, 0.6)]
    scn = scenarios.straight_road_test_world(n_others=1, multi_agent=True, primitives=prim, rules=[
        RuleMonitor.MakeRule("G !collision_corridor", -7.0, 0)])
    scn.params.REMOVE_AGENTS_OUT_OF_MAP = True
    scn.initial_state[1] = (30.0, -5.25, 0.0, 5.0)
    cfg = oracle_config(scn)
    acts = np.array([[0, 1]] * 12, np.uint8)       # agent 1 steers off the road, the ego keeps its lane
    o = orc.replay(cfg, scn.initial_state, FL(2), 20, acts)
    gone = np.flatnonzero(o["flags"][:o["steps"], 1] == 0)
    assert len(gone) and gone[0] > 0
    t = gone[0]
    assert o["rewards"][t, 1, 0] == -800.0 and np.all(o["rewards"][t + 1:o["steps"], 1] == 0.0)
    assert np.all(o["viol"][:, 1] == 0)             # it never got to see itself off the road: no rule penalty
    assert o["term_reward"][1, 0] == 0.0 and not o["terminal"][t]
    # the same sequence with the knob off: the agent stays, the rule fires, no out-of-map bonus branch
    scn.params.REMOVE_AGENTS_OUT_OF_MAP = False
    o2 = orc.replay(oracle_config(scn), scn.initial_state, FL(2), 20, acts)
    assert np.all(o2["flags"][:o2["steps"], 1] & ocw.PRESENT) and o2["viol"][o2["steps"] - 1, 1, 0] > 0
    assert o2["rewards"][t, 1, 0] == -800.0 - 7.0
    # the ego leaving the map ends the episode
    scn.params.REMOVE_AGENTS_OUT_OF_MAP = True
    o3 = orc.replay(oracle_config(scn), scn.initial_state, FL(2), 20, np.array([[1, 0]] * 12, np.uint8))
    k = o3["steps"]
    assert k < 12 and o3["terminal"][k - 1] and o3["flags"][k - 1, 0] == 0 and o3["rewards"][k - 1, 0, 0] == -800.0


def test_evaluate_rules_entry_matches_execute_labels(orc):
    """EvaluateRules() (:186-197) on Execute's successor reproduces the labels Execute saw."""
    scn = scenarios.merge_scenario(seed=4)
    cfg = oracle_config(scn)
    rng = np.random.default_rng(0)
    acts = rng.integers(0, 3, size=(6, 4)).astype(np.uint8)
    o = orc.replay(cfg, scn.initial_state, FL(4), 20, acts)
    for t in range(1, o["steps"]):
        e = orc.evaluate_rules(cfg, o["states"][t], o["flags"][t - 1], 20 - t - 1, o["q"][t - 1])
        ev = (o["flags"][t - 1, :4] & ocw.VALID) != 0
        assert np.array_equal(e["labels"][ev], o["labels"][t][ev])
        assert np.array_equal(e["q"][ev], o["q"][t][ev])


def test_oracle_reproduces_committed_golden_traces(orc):
    """tests/golden/cw_golden.npz freezes the FP64 oracle's traces (regression guard; these are
    the oracle's own outputs, not the reference's -- see tests/golden/make_cw_golden.py)."""
    import os
    from tests.golden.make_cw_golden import SCENARIOS
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cw_golden.npz"))
    for name, make in SCENARIOS.items():
        scn = make()
        cfg = oracle_config(scn)
        acts = g[name + "/actions"]
        for b in range(acts.shape[0]):
            o = orc.replay(cfg, scn.initial_state, FL(len(scn.agents)), scn.params.HORIZON, acts[b])
            k = int(g[name + "/steps"][b])
            assert o["steps"] == k
            for key in ["flags", "terminal", "q", "viol", "labels"]:
                assert np.array_equal(o[key][:k], g[name + "/" + key][b][:k]), (name, b, key)
            for key in ["states", "rewards", "frenet"]:
                np.testing.assert_allclose(o[key][:k], g[name + "/" + key][b][:k], rtol=1e-12, atol=1e-12)
