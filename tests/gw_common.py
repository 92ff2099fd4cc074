"""Shared helpers of the GridWorld tests: scenario/rule conversion between the product
ABI and the oracle, and random test-state generation."""
import numpy as np

from oracle import gw as ogw


def oracle_config_from_engine(eng):
    """Build the oracle's gwo_config from a planner_rules_mcts_b200.GridWorldEngine, so
    both sides run the same parameters and the same monitor tables."""
    p = eng.params
    rules = []
    for r in eng.rules:
        t = r.table
        tab = dict(aps=[eng_label(a) for a in t.aps], n_states=t.n_states, init=t.init,
                   trans=t.trans, accepting=t.accepting)
        rules.append(ogw.make_rule(tab, r.weight, r.priority, r.on_violation == "reset"))
    cfg = ogw.base_test_env_config(n_agents=p.num_other_agents + 1, terminal_depth=p.terminal_depth_,
                                   merging_point=p.merging_point, rules=rules)
    cfg.state_x_length = p.state_x_length
    cfg.ego_goal_reached_position = p.ego_goal_reached_position
    cfg.reward_vec_size = p.reward_vec_size
    cfg.n_actions = len(p.action_map)
    for k, d in enumerate(p.action_map):
        cfg.action_map[k] = d
    cfg.speed_deviation_weight = p.speed_deviation_weight
    cfg.acceleration_weight = p.acceleration_weight
    cfg.potential_weight = p.potential_weight
    return cfg


_LABEL = {"collision": ogw.LABEL_COLLISION, "merged_e": ogw.LABEL_MERGED_E,
          "merged_x": ogw.LABEL_MERGED_X, "in_direct_front_x": ogw.LABEL_IN_DIRECT_FRONT_X}


def eng_label(ap_name):
    return _LABEL[ap_name.split("#")[0]]


def random_states(rng, n, A, merging_point=8, spread=8, p_terminal=0.0):
    """n random but consistent GridWorld states: [n][A][4] agents, [n][A] terminal flags."""
    x = rng.integers(merging_point - spread, merging_point + 4, size=(n, A))
    last = rng.integers(-1, 2, size=(n, A))
    init_lane = rng.integers(1, 4, size=(n, A))
    lane = np.where(x >= merging_point, 0, init_lane)
    agents = np.stack([x, last, lane, init_lane], axis=-1).astype(np.int32)
    term = (rng.random((n, A)) < p_terminal).astype(np.uint8)
    term[:, 0] = 0
    return agents, term


def oracle_labels_to_words(mask, A):
    """oracle label bitmask (bit0 collision, bit1 merged_e, 2+j merged_x, 2+A+j front)
    -> the product's 3-word layout [ego bits, merged mask, front mask]."""
    mask = np.asarray(mask, np.uint32)
    ego = mask & 3
    merged = (mask >> 2) & ((1 << A) - 1)
    front = (mask >> (2 + A)) & ((1 << A) - 1)
    return np.stack([ego, merged, front], axis=-1).astype(np.uint32)
