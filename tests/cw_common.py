"""Shared helpers of the continuous-world tests: conversion of a product-side Scenario into
the oracle's configuration (same parameters, same monitor tables) and comparison policy."""
import numpy as np

from oracle import cw as ocw

# Tolerances (BASELINE.json north_star): kinematics and rewards within 1e-5 relative of the
# FP64 oracle; discrete outputs (labels, automaton states, violation counts, flags) bit-exact
# wherever the oracle's smallest threshold margin exceeds GUARD (fp32 rounding of a ~300 m
# coordinate is ~3e-5 m; positions accumulate a few ulps over the sub-steps).
RTOL = 1e-5
ATOL_POS = 2e-4      # metres (fp32 ulp of a 300 m coordinate is 3e-5)
GUARD = 5e-4


def oracle_config(scn, precision=0):
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
    c.precision = precision
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


def close(a, b, rtol=RTOL, atol=0.0):
    return np.allclose(np.asarray(a, np.float64), np.asarray(b, np.float64), rtol=rtol, atol=atol)


def compare_trace(tr, b, o, scn, min_exact_frac=0.0):
    """GPU trace `tr` (batch index b) vs oracle replay `o`.  Discrete outputs must be equal
    up to the first step whose oracle margin is below GUARD (after that the two runs may
    legitimately follow different branches); continuous outputs within tolerance on the same
    prefix.  Returns the number of steps compared."""
    k = o["steps"]
    safe = k
    for t in range(k):
        if o["margin"][t] < GUARD:
            safe = t
            break
    if safe == k:
        assert tr["steps_out"][b] == k
    A = len(scn.agents)
    for t in range(safe):
        assert np.array_equal(tr["flags_out"][b, t], o["flags"][t]), ("flags", b, t)
        assert tr["terminal_out"][b, t] == o["terminal"][t], ("terminal", b, t)
        assert np.array_equal(tr["labels_out"][b, t], o["labels"][t]), ("labels", b, t)
        assert np.array_equal(tr["q_out"][b, t], o["q"][t]), ("q", b, t)
        assert np.array_equal(tr["viol_out"][b, t], o["viol"][t]), ("viol", b, t)
        assert np.array_equal(tr["frenet_out"][b, t, :, 0], o["frenet"][t, :, 0]), ("lane", b, t)
        assert close(tr["agents_out"][b, t], o["states"][t], atol=ATOL_POS), ("state", b, t)
        assert close(tr["frenet_out"][b, t, :, 1:], o["frenet"][t, :, 1:], atol=ATOL_POS), ("frenet", b, t)
        assert close(tr["rewards_out"][b, t], o["rewards"][t], rtol=1e-5, atol=2e-4), ("rewards", b, t)
    if safe == k:
        assert close(tr["term_reward_out"][b], o["term_reward"], atol=1e-6)
    return safe


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
    out = eng.rollout(eng.initial_batch(), 512, seed=5, max_steps=20)
    ref = orc.rollout(cfg, scn.initial_state, flags0, scn.params.HORIZON, 5, 0, 512, 20, scn.params.DISCOUNT_FACTOR)
    assert abs(int(out["agent_steps"][0]) - ref["agent_steps"]) <= 0.02 * ref["agent_steps"]
    assert np.allclose(out["returns_sum"][0], ref["returns_sum"], rtol=2e-2, atol=1.0)
