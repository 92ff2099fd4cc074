"""Shared helpers of the continuous-world tests: conversion of a product-side Scenario into
the oracle's configuration (same parameters, same monitor tables) and comparison policy."""
import numpy as np

from oracle import cw as ocw

# Parity bar (BASELINE.json north_star: discrete outputs bit-exact, floats within 1e-5
# relative).  The device computes the continuous world in FP64 like the reference, so the
# tests ask for much more than north_star does:
#   * labels, automaton states, violation counts, flags, terminal flags, lane indices, step
#     counts: EQUAL on every trace and every rollout -- no exclusion of any kind;
#   * kinematics, Frenet coordinates, rewards, returns: within RTOL = 1e-9 relative (measured:
#     ~1e-13; the device uses closed forms of the Euler sub-steps, fused multiply-adds and its
#     own sine / cosine, the oracle the sub-stepped loop and libm), ATOL for values that are
#     exactly 0 in one of the two (lateral offsets on the centre line).
RTOL = 1e-9
ATOL = 1e-9


def oracle_config(scn):
    """cwo_config from a planner_rules_mcts_b200.Scenario."""
    p = scn.params
    c = ocw.Config()
    c.collision_weight = p.COLLISION_WEIGHT
    c.out_of_map_weight = p.OUT_OF_MAP_WEIGHT
    c.potential_weight = p.POTENTIAL_WEIGHT
    c.acceleration_weight = p.ACCELERATION_WEIGHT
    c.radial_acceleration_weight = p.RADIAL_ACCELERATION_WEIGHT
    c.desired_velocity_weight = p.DESIRED_VELOCITY_WEIGHT
    c.lane_center_weight = p.LANE_CENTER_WEIGHT
    c.prediction_time_span = p.PREDICTION_TIME_SPAN
    c.desired_velocity = p.DESIRED_VELOCITY
    c.discount_factor = p.DISCOUNT_FACTOR
    c.goal_reward = p.GOAL_REWARD
    c.reward_vector_size = p.REWARD_VECTOR_SIZE
    c.horizon = p.HORIZON
    c.use_rule_reward_for_ego_only = 1 if p.USE_RULE_REWARD_FOR_EGO_ONLY else 0
    c.remove_agents_out_of_map = 1 if p.REMOVE_AGENTS_OUT_OF_MAP else 0
    c.wheel_base = p.WHEEL_BASE
    c.integration_time_delta = p.INTEGRATION_TIME_DELTA
    c.sd_reaction_time = p.SAFE_DISTANCE_REACTION_TIME
    c.sd_max_decel_rear = p.SAFE_DISTANCE_MAX_DECEL_REAR
    c.sd_max_decel_front = p.SAFE_DISTANCE_MAX_DECEL_FRONT
    c.near_distance = p.NEAR_DISTANCE
    c.n_agents = len(scn.agents)
    c.n_mcts_agents = scn.n_mcts_agents
    c.n_lanes = len(scn.lanes)
    road = np.asarray(scn.road, np.float64)
    c.n_road = road.shape[0]
    c.n_rules = len(scn.rules)
    for i, a in enumerate(scn.agents):
        o = c.agents[i]
        o.front, o.rear, o.half_width = a.front, a.rear, a.half_width
        o.has_goal = 0 if a.goal is None else 1
        g = [0.0, 0.0, 0.0, 0.0, -1e30, 1e30]
        if a.goal is not None:
            g[:len(a.goal)] = a.goal
        for k in range(6):
            o.goal[k] = g[k]
        o.n_prims = len(a.primitives)
        for k, (acc, delta) in enumerate(a.primitives):
            o.prim_acc[k] = acc
            o.prim_delta[k] = delta
    for i, l in enumerate(scn.lanes):
        o = c.lanes[i]
        pts = np.asarray(l.points, np.float64)
        o.n_pts = pts.shape[0]
        o.successor = l.successor
        o.flags = 1 if l.rightmost else 0
        o.half_width = l.half_width
        o.s_point = l.s_point
        for k in range(pts.shape[0]):
            o.pts[k][0] = pts[k, 0]
            o.pts[k][1] = pts[k, 1]
    for k in range(road.shape[0]):
        c.road[k][0] = road[k, 0]
        c.road[k][1] = road[k, 1]
    from planner_rules_mcts_b200.continuous import CW_LABELS
    for k, r in enumerate(scn.rules):
        t = r.table
        o = c.rules[k]
        n_states, rows, accepting = r.packed_table()   # with the "does not apply" state, as uploaded
        o.n_aps, o.n_states, o.init_state = t.n_aps, n_states, t.init
        o.priority = r.priority
        o.reset_on_violation = 1 if r.on_violation == "reset" else 0
        o.weight = r.weight
        for j, name in enumerate(t.aps):
            o.ap_label[j] = CW_LABELS[name.split("#")[0]]
        stride = 1 << t.n_aps
        for q, row in enumerate(rows):
            for l, nxt in enumerate(row):
                o.trans[q * stride + l] = nxt
        for q, acc in enumerate(accepting):
            o.accepting[q] = acc
    return c


def close(a, b, rtol=RTOL, atol=ATOL):
    return np.allclose(np.asarray(a, np.float64), np.asarray(b, np.float64), rtol=rtol, atol=atol)


def assert_close(a, b, what, rtol=RTOL, atol=ATOL):
    np.testing.assert_allclose(np.asarray(a, np.float64), np.asarray(b, np.float64), rtol=rtol, atol=atol,
                               err_msg=str(what))


def compare_trace(tr, b, o, scn):
    """Trace `tr` (batch index b; brm_cw_trace layout) vs oracle replay `o`: every discrete output
    equal on every step, continuous outputs within RTOL.  Returns the number of steps."""
    k = o["steps"]
    assert tr["steps_out"][b] == k, ("steps", b, tr["steps_out"][b], k)
    for key_t, key_o in (("flags_out", "flags"), ("terminal_out", "terminal"), ("labels_out", "labels"),
                         ("q_out", "q"), ("viol_out", "viol")):
        assert np.array_equal(tr[key_t][b, :k], o[key_o][:k]), (key_t, b)
    assert np.array_equal(tr["frenet_out"][b, :k, :, 0], o["frenet"][:k, :, 0]), ("lane", b)
    assert_close(tr["agents_out"][b, :k], o["states"][:k], ("state", b))
    assert_close(tr["frenet_out"][b, :k, :, 1:], o["frenet"][:k, :, 1:], ("frenet", b))
    assert_close(tr["rewards_out"][b, :k], o["rewards"][:k], ("rewards", b))
    assert_close(tr["term_reward_out"][b], o["term_reward"], ("term_reward", b))
    return k


def compare_rollouts(out, ref, what=""):
    """Per-rollout outputs of brm_cw_rollout (`ro_*`) vs cwo_rollout_detail: all discrete outputs
    equal on ALL rollouts, final states and returns within RTOL."""
    assert np.array_equal(out["ro_steps"], ref["steps"]), (what, "steps",
                                                            int((out["ro_steps"] != ref["steps"]).sum()))
    assert np.array_equal(out["ro_viol"], ref["viol"]), (what, "viol")
    assert np.array_equal(out["ro_q"], ref["q"]), (what, "q")
    assert np.array_equal(out["ro_flags"], ref["flags"]), (what, "flags")
    assert_close(out["ro_agents"], ref["states"], (what, "states"))
    assert_close(out["ro_returns"], ref["returns"], (what, "returns"))


def smoke_cw():
    """One small continuous-world invocation on cuda:0 checked against the oracle."""
    import planner_rules_mcts_b200 as brm
    from planner_rules_mcts_b200 import scenarios

    scn = scenarios.merge_scenario(seed=3)
    eng = brm.RulesMctsEngine([scn], device=0)
    cfg = oracle_config(scn)
    orc = ocw.CwOracle()
    rng = np.random.default_rng(0)
    B, T = 8, 12
    acts = rng.integers(0, 3, size=(B, T, scn.n_mcts_agents)).astype(np.uint8)
    batch = eng.make_batch(np.broadcast_to(scn.initial_state, (B,) + scn.initial_state.shape))
    tr = eng.replay(batch, acts)
    flags0 = np.full(len(scn.agents), 3, np.uint8)
    compared = 0
    for b in range(B):
        o = orc.replay(cfg, scn.initial_state, flags0, scn.params.HORIZON, acts[b])
        compared += compare_trace(tr, b, o, scn)
    assert compared > 0
    out = eng.rollout(eng.initial_batch(), 512, seed=5, max_steps=20, detail=True)
    ref = orc.rollout_detail(cfg, scn.initial_state, flags0, scn.params.HORIZON, 5, 0, 512, 20,
                             scn.params.DISCOUNT_FACTOR)
    compare_rollouts(out, ref, "smoke")
