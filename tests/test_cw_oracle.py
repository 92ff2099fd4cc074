"""Continuous-world CPU oracle (oracle/cw_oracle.cpp) against the only numeric pins the
reference holds for the BARK-backed state: test/single_agent_test.cpp:76-174 (rewards 0 /
-800 / -1000, terminal flags, goal reached by repeating action 0) and
test/multi_agent_test.cpp:51-140 (3 -> 1 actions after a collision, frozen positions).
Everything else about this oracle is "parity unpinned" (see oracle/cw_oracle.h)."""
import numpy as np
import pytest

from oracle import cw as ocw
from planner_rules_mcts_b200 import scenarios
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
        dt = float(np.float32(p.PREDICTION_TIME_SPAN))
        a = (v1 - v0) / dt
        cost = p.ACCELERATION_WEIGHT * a * a * dt
        a_lat = (th1 - th0) / dt * (0.5 * a * dt + v0)
        cost += p.RADIAL_ACCELERATION_WEIGHT * a_lat * a_lat * dt
        cost += p.DESIRED_VELOCITY_WEIGHT * abs(v0 - p.DESIRED_VELOCITY) * dt
        cost += p.LANE_CENTER_WEIGHT * abs(y1 + 1.75) * dt
        pot = lambda v: -p.POTENTIAL_WEIGHT * dt * abs(v - p.DESIRED_VELOCITY)
        want = cost + float(np.float32(p.DISCOUNT_FACTOR)) * pot(v1) - pot(v0)
        assert abs(o["rewards"][0, 0, 0] - want) < 1e-9 * max(1.0, abs(want))
        assert abs(o["frenet"][0, 0, 2] - (y1 + 1.75)) < 1e-9 and o["frenet"][0, 0, 0] == 0


def test_euler_substeps(orc):
    # 15 explicit Euler sub-steps of 0.02 s (bark BehaviorMPContinuousActions, un-vendored)
    scn = scenarios.straight_road_test_world(n_others=0, primitives=[(2.0, 0.05)], rules=[])
    cfg = oracle_config(scn)
    o = orc.replay(cfg, scn.initial_state, FL(1), 20, np.array([[0]], np.uint8))
    x, y, th, v = scn.initial_state[0].astype(np.float64)
    k = np.tan(0.05) / 2.7
    k = float(np.float32(k))
    h = float(np.float32(0.02))
    for n in range(15):
        hh = float(np.float32(0.3 - 14 * 0.02)) if n == 14 else h
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


def test_float_instantiation_tracks_double(orc):
    scn = scenarios.highway_scenario()
    d = orc.rollout_detail(oracle_config(scn, 0), scn.initial_state, FL(4), 20, 7, 0, 300, 20, 0.9)
    f = orc.rollout_detail(oracle_config(scn, 1), scn.initial_state, FL(4), 20, 7, 0, 300, 20, 0.9)
    same = (d["steps"] == f["steps"]) & (d["viol"] == f["viol"]).all(axis=(1, 2))
    assert same.mean() > 0.98
    np.testing.assert_allclose(f["states"][same], d["states"][same], rtol=1e-4, atol=1e-2)


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
