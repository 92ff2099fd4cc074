"""TEST INFRASTRUCTURE -- ctypes access to the GridWorld CPU oracles.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  See oracle/gw_oracle.h for the two implementations
(`ref` = the reference's own sources under oracle/_ref, `port` = oracle/gw_oracle.c).
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

MAX_AGENTS, MAX_APS, MAX_STATES, MAX_RULES, MAX_ACTIONS = 8, 4, 16, 8, 8
VIOLATION = 0xFF
LABEL_COLLISION, LABEL_MERGED_E, LABEL_MERGED_X, LABEL_IN_DIRECT_FRONT_X = 0, 1, 2, 3


class Rule(C.Structure):
    _fields_ = [
        ("n_aps", C.c_int32),
        ("n_states", C.c_int32),
        ("init_state", C.c_int32),
        ("priority", C.c_int32),
        ("reset_on_violation", C.c_int32),
        ("weight", C.c_float),
        ("ap_label", C.c_int32 * MAX_APS),
        ("trans", C.c_uint8 * (MAX_STATES * 16)),
        ("accepting", C.c_uint8 * MAX_STATES),
    ]


class Config(C.Structure):
    _fields_ = [
        ("n_agents", C.c_int32),
        ("state_x_length", C.c_int32),
        ("ego_goal_reached_position", C.c_int32),
        ("terminal_depth", C.c_int32),
        ("merging_point", C.c_int32),
        ("reward_vec_size", C.c_int32),
        ("n_actions", C.c_int32),
        ("action_map", C.c_int32 * MAX_ACTIONS),
        ("speed_deviation_weight", C.c_double),
        ("acceleration_weight", C.c_double),
        ("potential_weight", C.c_double),
        ("n_rules", C.c_int32),
        ("_pad", C.c_int32),
        ("rules", Rule * MAX_RULES),
    ]


# Hand-derived monitor tables, SURVEY.md Appendix B (normative for the oracle).
# `G !collision`: one state, p -> violation.
G_NOT_COLLISION = dict(aps=[LABEL_COLLISION], n_states=1, init=0,
                       trans=[[0, VIOLATION]], accepting=[1])
# ZIP_FORMULA (/root/reference/gridworld/base_test_env.cpp:16-19), APs (f, e, x) =
# (in_direct_front_x#0, merged_e, merged_x#0); letter = f<<2 | e<<1 | x.
V_ = VIOLATION
ZIP = dict(aps=[LABEL_IN_DIRECT_FRONT_X, LABEL_MERGED_E, LABEL_MERGED_X], n_states=4, init=0,
           trans=[[3, 3, 3, 3, 1, 1, 3, 3],
                  [3, 1, 2, 2, 1, 1, 2, V_],
                  [2, 2, 2, 2, 2, 2, 2, V_],
                  [3, 3, 3, 3, 3, 3, 3, 3]],
           accepting=[1, 1, 1, 1])


def make_rule(table, weight, priority, reset_on_violation=True):
    r = Rule()
    r.n_aps = len(table["aps"])
    r.n_states = table["n_states"]
    r.init_state = table["init"]
    r.priority = priority
    r.reset_on_violation = 1 if reset_on_violation else 0
    r.weight = weight
    for k, a in enumerate(table["aps"]):
        r.ap_label[k] = a
    stride = 1 << r.n_aps
    for s, row in enumerate(table["trans"]):
        assert len(row) == stride
        for l, nxt in enumerate(row):
            r.trans[s * stride + l] = nxt
    for s, a in enumerate(table["accepting"]):
        r.accepting[s] = a
    return r


def base_test_env_config(n_agents=3, terminal_depth=8, merging_point=8, rules=None,
                         reset_on_violation=True):
    """Scenario G: /root/reference/gridworld/base_test_env.cpp:45-102."""
    cfg = Config()
    cfg.n_agents = n_agents
    cfg.state_x_length = 1000
    cfg.ego_goal_reached_position = 1000
    cfg.terminal_depth = terminal_depth
    cfg.merging_point = merging_point
    cfg.reward_vec_size = 3
    cfg.n_actions = 3
    for k, d in enumerate((1, 0, -1)):  # FORWARD, WAIT, BACKWARD (:83-85)
        cfg.action_map[k] = d
    cfg.speed_deviation_weight = 5.0
    cfg.acceleration_weight = 1.5
    cfg.potential_weight = 0.0
    if rules is None:
        rules = [make_rule(G_NOT_COLLISION, -1.0, 0, reset_on_violation),
                 make_rule(ZIP, -1.0, 1, reset_on_violation)]
    cfg.n_rules = len(rules)
    for k, r in enumerate(rules):
        cfg.rules[k] = r
    return cfg


def base_test_env_state(merging_point=8):
    """Initial agent states of BaseTestEnv (base_test_env.cpp:29-43): agents 0,1 share
    lane 1, agent 2 drives on lane 3 (AgentState() draws lanes from a global counter:
    gridworld/common.hpp:48-53), last_action = FORWARD."""
    agents = np.array([[merging_point - 6, 1, 1, 1],
                       [merging_point - 2, 1, 1, 1],
                       [merging_point - 8, 1, 3, 3]], dtype=np.int32)
    term = np.zeros(3, dtype=np.uint8)
    return agents, term


def _ptr(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


class GwOracle:
    """impl = 'port' (oracle/gw_oracle.c) or 'ref' (the reference's own sources)."""

    def __init__(self, impl="port"):
        self.impl = impl
        if impl == "port":
            path = os.path.join(HERE, "libgw_oracle.so")
            self.prefix = "gwo_"
        else:
            path = os.path.join(HERE, "_ref", "libgw_ref.so")
            self.prefix = "gwref_"
        if not os.path.exists(path):
            build()
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = C.CDLL(path)
        self.lib[self.prefix + "rollout"].restype = C.c_int32
        self.lib[self.prefix + "replay"].restype = C.c_int32

    def rules_per_agent(self, cfg):
        return int(self.lib[self.prefix + "rules_per_agent"](C.byref(cfg)))

    def replay(self, cfg, agents0, term0, depth0, actions, q0=None):
        A, V = cfg.n_agents, cfg.reward_vec_size
        R = self.rules_per_agent(cfg)
        actions = np.ascontiguousarray(actions, dtype=np.uint8)
        T = actions.shape[0]
        agents0 = np.ascontiguousarray(agents0, dtype=np.int32)
        term0 = np.ascontiguousarray(term0, dtype=np.uint8)
        out = dict(
            agents=np.zeros((T, A, 4), np.int32), term=np.zeros((T, A), np.uint8),
            q=np.zeros((T, A, R), np.uint8), viol=np.zeros((T, A, R), np.uint32),
            rewards=np.zeros((T, A, V), np.float64), labels=np.zeros((T, A), np.uint32),
            term_reward=np.zeros((A, V), np.float64))
        q0a = None if q0 is None else np.ascontiguousarray(q0, dtype=np.uint8)
        n = self.lib[self.prefix + "replay"](
            C.byref(cfg), _ptr(agents0, C.c_int32), _ptr(term0, C.c_uint8), C.c_int32(depth0),
            _ptr(q0a, C.c_uint8), _ptr(actions, C.c_uint8), C.c_int32(T),
            _ptr(out["agents"], C.c_int32), _ptr(out["term"], C.c_uint8),
            _ptr(out["q"], C.c_uint8), _ptr(out["viol"], C.c_uint32),
            _ptr(out["rewards"], C.c_double), _ptr(out["labels"], C.c_uint32),
            _ptr(out["term_reward"], C.c_double))
        out["steps"] = int(n)
        return out

    def rollout(self, cfg, agents0, term0, depth0, seed, rollout_begin, n_rollouts, max_steps,
                gamma, first_discount_exp=1, n_threads=1, q0=None):
        A, V = cfg.n_agents, cfg.reward_vec_size
        R = self.rules_per_agent(cfg)
        agents0 = np.ascontiguousarray(agents0, dtype=np.int32)
        term0 = np.ascontiguousarray(term0, dtype=np.uint8)
        ret = np.zeros((A, V), np.float64)
        viol = np.zeros((A, R), np.uint64)
        steps = C.c_uint64(0)
        q0a = None if q0 is None else np.ascontiguousarray(q0, dtype=np.uint8)
        self.lib[self.prefix + "rollout"](
            C.byref(cfg), _ptr(agents0, C.c_int32), _ptr(term0, C.c_uint8), C.c_int32(depth0),
            _ptr(q0a, C.c_uint8), C.c_uint64(seed), C.c_uint64(rollout_begin),
            C.c_uint64(n_rollouts), C.c_int32(max_steps), C.c_double(gamma),
            C.c_int32(first_discount_exp), C.c_int32(n_threads), _ptr(ret, C.c_double),
            _ptr(viol, C.c_uint64), C.byref(steps))
        return dict(returns_sum=ret, viol_sum=viol, agent_steps=int(steps.value))

    def rollout_detail(self, cfg, agents0, term0, depth0, seed, rollout_begin, n_rollouts,
                       max_steps, gamma, first_discount_exp=1, q0=None):
        assert self.impl == "port"
        A, V = cfg.n_agents, cfg.reward_vec_size
        R = self.rules_per_agent(cfg)
        agents0 = np.ascontiguousarray(agents0, dtype=np.int32)
        term0 = np.ascontiguousarray(term0, dtype=np.uint8)
        n = n_rollouts
        out = dict(returns=np.zeros((n, A, V), np.float64), viol=np.zeros((n, A, R), np.uint32),
                   q=np.zeros((n, A, R), np.uint8), agents=np.zeros((n, A, 4), np.int32),
                   steps=np.zeros(n, np.int32))
        q0a = None if q0 is None else np.ascontiguousarray(q0, dtype=np.uint8)
        self.lib.gwo_rollout_detail(
            C.byref(cfg), _ptr(agents0, C.c_int32), _ptr(term0, C.c_uint8), C.c_int32(depth0),
            _ptr(q0a, C.c_uint8), C.c_uint64(seed), C.c_uint64(rollout_begin), C.c_uint64(n),
            C.c_int32(max_steps), C.c_double(gamma), C.c_int32(first_discount_exp),
            _ptr(out["returns"], C.c_double), _ptr(out["viol"], C.c_uint32),
            _ptr(out["q"], C.c_uint8), _ptr(out["agents"], C.c_int32),
            _ptr(out["steps"], C.c_int32))
        return out

    def philox(self, ctr, key):
        assert self.impl == "port"
        c = (C.c_uint32 * 4)(*ctr)
        k = (C.c_uint32 * 2)(*key)
        o = (C.c_uint32 * 4)()
        self.lib.gwo_philox4x32_10(c, k, o)
        return list(o)

    def at_position(self, x, last, pos):
        return bool(self.lib[self.prefix + "at_position"](x, last, pos))

    def label_collision(self, mp, ex, el, elane, ox, ol, olane):
        return bool(self.lib[self.prefix + "label_collision"](mp, ex, el, elane, ox, ol, olane))


def build(ref=True):
    """Compile the oracles (make -C oracle).  `ref` needs /root/reference; on a GPU box
    the prebuilt oracle/_ref/libgw_ref.so that travelled with the snapshot is used."""
    targets = ["oracle"] + (["ref"] if ref else [])
    subprocess.run(["make", "-C", HERE] + targets, check=True, stdout=subprocess.DEVNULL)
