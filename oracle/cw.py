"""TEST INFRASTRUCTURE -- ctypes access to the continuous-world CPU oracle
(oracle/cw_oracle.cpp; see oracle/cw_oracle.h: "parity unpinned" for the un-vendored
dependencies).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
MAX_AGENTS, MAX_LANES, MAX_PTS, MAX_ROAD, MAX_PRIMS, MAX_RULES, MAX_SLOTS = 16, 8, 64, 32, 8, 8, 32
MAX_APS, MAX_STATES = 4, 16
PRESENT, VALID = 1, 2


class Rule(C.Structure):
    _fields_ = [("n_aps", C.c_int32), ("n_states", C.c_int32), ("init_state", C.c_int32),
                ("priority", C.c_int32), ("reset_on_violation", C.c_int32), ("weight", C.c_float),
                ("ap_label", C.c_int32 * MAX_APS), ("trans", C.c_uint8 * (MAX_STATES * 16)),
                ("accepting", C.c_uint8 * MAX_STATES)]


class Agent(C.Structure):
    _fields_ = [("front", C.c_double), ("rear", C.c_double), ("half_width", C.c_double),
                ("has_goal", C.c_int32), ("n_prims", C.c_int32), ("goal", C.c_double * 6),
                ("prim_acc", C.c_double * MAX_PRIMS), ("prim_delta", C.c_double * MAX_PRIMS)]


class Lane(C.Structure):
    _fields_ = [("n_pts", C.c_int32), ("successor", C.c_int32), ("flags", C.c_int32), ("_pad", C.c_int32),
                ("half_width", C.c_double), ("s_point", C.c_double), ("pts", (C.c_double * 2) * MAX_PTS)]


class Config(C.Structure):
    _fields_ = [(k, C.c_double) for k in (
        "collision_weight", "out_of_map_weight", "potential_weight", "acceleration_weight",
        "radial_acceleration_weight", "desired_velocity_weight", "lane_center_weight",
        "prediction_time_span", "desired_velocity", "discount_factor", "goal_reward")] + [
        ("reward_vector_size", C.c_int32), ("horizon", C.c_int32),
        ("use_rule_reward_for_ego_only", C.c_int32), ("remove_agents_out_of_map", C.c_int32),
        ("wheel_base", C.c_double), ("integration_time_delta", C.c_double),
        ("sd_reaction_time", C.c_double), ("sd_max_decel_rear", C.c_double),
        ("sd_max_decel_front", C.c_double), ("near_distance", C.c_double),
        ("n_agents", C.c_int32), ("n_mcts_agents", C.c_int32), ("n_lanes", C.c_int32),
        ("n_road", C.c_int32), ("n_rules", C.c_int32), ("_pad", C.c_int32),
        ("agents", Agent * MAX_AGENTS), ("lanes", Lane * MAX_LANES),
        ("road", (C.c_double * 2) * MAX_ROAD), ("rules", Rule * MAX_RULES)]


def _ptr(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


class CwOracle:
    def __init__(self):
        path = os.path.join(HERE, "libcw_oracle.so")
        if not os.path.exists(path):
            subprocess.run(["make", "-C", HERE, "oracle"], check=True, stdout=subprocess.DEVNULL)
        self.lib = C.CDLL(path)

    def rules_per_agent(self, cfg):
        return int(self.lib.cwo_rules_per_agent(C.byref(cfg)))

    def check_terminal(self, cfg, states, flags, horizon):
        states = np.ascontiguousarray(states, np.float64)
        flags = np.ascontiguousarray(flags, np.uint8)
        return bool(self.lib.cwo_check_terminal(C.byref(cfg), _ptr(states, C.c_double), _ptr(flags, C.c_uint8),
                                                C.c_int32(horizon)))

    def replay(self, cfg, states0, flags0, horizon0, actions, q0=None):
        A, M, V = cfg.n_agents, cfg.n_mcts_agents, cfg.reward_vector_size
        R = self.rules_per_agent(cfg)
        actions = np.ascontiguousarray(actions, np.uint8)
        T = actions.shape[0]
        states0 = np.ascontiguousarray(states0, np.float64)
        flags0 = np.ascontiguousarray(flags0, np.uint8)
        q0a = None if q0 is None else np.ascontiguousarray(q0, np.uint8)
        o = dict(states=np.zeros((T, A, 4)), flags=np.zeros((T, A), np.uint8), terminal=np.zeros(T, np.uint8),
                 q=np.zeros((T, M, R), np.uint8), viol=np.zeros((T, M, R), np.uint32),
                 rewards=np.zeros((T, M, V)), labels=np.zeros((T, M, 4), np.uint32),
                 frenet=np.zeros((T, A, 3)), margin=np.full(T, np.inf), term_reward=np.zeros((M, V)))
        self.lib.cwo_replay.restype = C.c_int32
        n = self.lib.cwo_replay(
            C.byref(cfg), _ptr(states0, C.c_double), _ptr(flags0, C.c_uint8), C.c_int32(horizon0),
            _ptr(q0a, C.c_uint8), _ptr(actions, C.c_uint8), C.c_int32(T), _ptr(o["states"], C.c_double),
            _ptr(o["flags"], C.c_uint8), _ptr(o["terminal"], C.c_uint8), _ptr(o["q"], C.c_uint8),
            _ptr(o["viol"], C.c_uint32), _ptr(o["rewards"], C.c_double), _ptr(o["labels"], C.c_uint32),
            _ptr(o["frenet"], C.c_double), _ptr(o["margin"], C.c_double), _ptr(o["term_reward"], C.c_double))
        o["steps"] = int(n)
        return o

    def rollout(self, cfg, states0, flags0, horizon0, seed, rollout_begin, n_rollouts, max_steps, gamma,
                first_discount_exp=1, n_threads=1, q0=None, coop_factor=0.0):
        M, V = cfg.n_mcts_agents, cfg.reward_vector_size
        R = self.rules_per_agent(cfg)
        states0 = np.ascontiguousarray(states0, np.float64)
        flags0 = np.ascontiguousarray(flags0, np.uint8)
        q0a = None if q0 is None else np.ascontiguousarray(q0, np.uint8)
        ret = np.zeros((M, V))
        viol = np.zeros((M, R), np.uint64)
        steps = C.c_uint64(0)
        self.lib.cwo_rollout(
            C.byref(cfg), _ptr(states0, C.c_double), _ptr(flags0, C.c_uint8), C.c_int32(horizon0),
            _ptr(q0a, C.c_uint8), C.c_uint64(seed), C.c_uint64(rollout_begin), C.c_uint64(n_rollouts),
            C.c_int32(max_steps), C.c_double(gamma), C.c_int32(first_discount_exp), C.c_double(coop_factor),
            C.c_int32(n_threads), _ptr(ret, C.c_double), _ptr(viol, C.c_uint64), C.byref(steps))
        return dict(returns_sum=ret, viol_sum=viol, agent_steps=int(steps.value))

    def evaluate_rules(self, cfg, states, flags, horizon, q):
        """MvmctsState::EvaluateRules() on one state -> dict(q, viol, rewards, labels)."""
        M, V = cfg.n_mcts_agents, cfg.reward_vector_size
        R = self.rules_per_agent(cfg)
        states = np.ascontiguousarray(states, np.float64)
        flags = np.ascontiguousarray(flags, np.uint8)
        q = np.ascontiguousarray(q, np.uint8)
        o = dict(q=np.zeros((M, R), np.uint8), viol=np.zeros((M, R), np.uint32), rewards=np.zeros((M, V)),
                 labels=np.zeros((M, 4), np.uint32))
        self.lib.cwo_evaluate_rules(C.byref(cfg), _ptr(states, C.c_double), _ptr(flags, C.c_uint8),
                                    C.c_int32(horizon), _ptr(q, C.c_uint8), _ptr(o["q"], C.c_uint8),
                                    _ptr(o["viol"], C.c_uint32), _ptr(o["rewards"], C.c_double),
                                    _ptr(o["labels"], C.c_uint32))
        return o

    def coop_mix(self, returns_sum, coop_factor):
        r = np.ascontiguousarray(returns_sum, np.float64).copy()
        self.lib.cwo_coop_mix.restype = None
        self.lib.cwo_coop_mix(_ptr(r, C.c_double), C.c_int32(r.shape[0]), C.c_int32(r.shape[1]),
                              C.c_double(coop_factor))
        return r

    def rollout_detail(self, cfg, states0, flags0, horizon0, seed, rollout_begin, n_rollouts, max_steps,
                       gamma, first_discount_exp=1, q0=None):
        A, M, V = cfg.n_agents, cfg.n_mcts_agents, cfg.reward_vector_size
        R = self.rules_per_agent(cfg)
        n = n_rollouts
        states0 = np.ascontiguousarray(states0, np.float64)
        flags0 = np.ascontiguousarray(flags0, np.uint8)
        q0a = None if q0 is None else np.ascontiguousarray(q0, np.uint8)
        o = dict(returns=np.zeros((n, M, V)), viol=np.zeros((n, M, R), np.uint32),
                 q=np.zeros((n, M, R), np.uint8), states=np.zeros((n, A, 4)), flags=np.zeros((n, A), np.uint8),
                 steps=np.zeros(n, np.int32), margin=np.zeros(n))
        self.lib.cwo_rollout_detail(
            C.byref(cfg), _ptr(states0, C.c_double), _ptr(flags0, C.c_uint8), C.c_int32(horizon0),
            _ptr(q0a, C.c_uint8), C.c_uint64(seed), C.c_uint64(rollout_begin), C.c_uint64(n),
            C.c_int32(max_steps), C.c_double(gamma), C.c_int32(first_discount_exp),
            _ptr(o["returns"], C.c_double), _ptr(o["viol"], C.c_uint32), _ptr(o["q"], C.c_uint8),
            _ptr(o["states"], C.c_double), _ptr(o["flags"], C.c_uint8), _ptr(o["steps"], C.c_int32),
            _ptr(o["margin"], C.c_double))
        return o
