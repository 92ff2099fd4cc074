"""Host mirror of the reference's BARK-backed state plugin `MvmctsState`
(/root/reference/src/rules_mcts_state.{hpp,cpp}) on top of the C ABI.

`RulesMctsStateParameters` keeps the reference's parameter names and defaults
(src/rules_mcts_state_parameters.cpp:36-78); `Scenario` is the self-contained world model
that replaces the BARK `ObservedWorld` (agents with rectangle shapes, motion primitives and
goals; lane centre lines; one road polygon); `RulesMctsEngine` is the batched surface
(Execute / GetNumActions / IsTerminal / GetTerminalReward / rollouts) and `MvmctsState`
the single-state view with the StateInterface member names.  All arithmetic runs in
libbrm.so on the GPU.
"""
import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from . import _abi
from .ltl import RuleMonitor

CW_LABELS = {
    "collision_ego": _abi.CW_LABEL_COLLISION_EGO,
    "collision_corridor": _abi.CW_LABEL_COLLISION_CORRIDOR,
    "goal_reached": _abi.CW_LABEL_GOAL_REACHED,
    "safe_distance_front": _abi.CW_LABEL_SAFE_DISTANCE_FRONT,
    "merged_e": _abi.CW_LABEL_MERGED_E,
    "rightmost_lane": _abi.CW_LABEL_RIGHTMOST_LANE,
    "on_road": _abi.CW_LABEL_ON_ROAD,
    "merged_x": _abi.CW_LABEL_MERGED_X,
    "in_direct_front_x": _abi.CW_LABEL_IN_DIRECT_FRONT_X,
    "near_x": _abi.CW_LABEL_NEAR_X,
}

PRESENT, VALID = _abi.CW_PRESENT, _abi.CW_VALID
BIG = 1.0e30


@dataclass
class RulesMctsStateParameters:
    """src/rules_mcts_state_parameters.hpp:33-46 with the defaults of the ParamsPtr
    constructor (src/rules_mcts_state_parameters.cpp:36-78), plus the constants the hot
    path reads from bark's parameter server (dynamics, label functions)."""
    COLLISION_WEIGHT: float = 0.0
    OUT_OF_MAP_WEIGHT: float = -800.0
    POTENTIAL_WEIGHT: float = 32.0
    ACCELERATION_WEIGHT: float = 0.0
    RADIAL_ACCELERATION_WEIGHT: float = 0.0
    DESIRED_VELOCITY_WEIGHT: float = -5.0
    LANE_CENTER_WEIGHT: float = 0.0
    REWARD_VECTOR_SIZE: int = 1
    PREDICTION_TIME_SPAN: float = 0.3
    DESIRED_VELOCITY: float = 10.0
    HORIZON: int = 30
    DISCOUNT_FACTOR: float = 0.9
    GOAL_REWARD: float = 0.0
    USE_RULE_REWARD_FOR_EGO_ONLY: bool = False
    # bark "World::remove_agents_out_of_map": Predict removes agents whose shape left the road
    # polygon; an MCTS agent that vanished earns OUT_OF_MAP_WEIGHT (src/rules_mcts_state.cpp:121-125)
    REMOVE_AGENTS_OUT_OF_MAP: bool = False
    # bark (un-vendored): SingleTrackModel wheel base, BehaviorMotionPrimitives integration step
    WHEEL_BASE: float = 2.7
    INTEGRATION_TIME_DELTA: float = 0.02
    # label functions
    SAFE_DISTANCE_REACTION_TIME: float = 1.0
    SAFE_DISTANCE_MAX_DECEL_REAR: float = 8.0
    SAFE_DISTANCE_MAX_DECEL_FRONT: float = 8.0
    NEAR_DISTANCE: float = 8.0

    def to_abi(self):
        p = _abi.CwParams()
        p.collision_weight = self.COLLISION_WEIGHT
        p.out_of_map_weight = self.OUT_OF_MAP_WEIGHT
        p.potential_weight = self.POTENTIAL_WEIGHT
        p.acceleration_weight = self.ACCELERATION_WEIGHT
        p.radial_acceleration_weight = self.RADIAL_ACCELERATION_WEIGHT
        p.desired_velocity_weight = self.DESIRED_VELOCITY_WEIGHT
        p.lane_center_weight = self.LANE_CENTER_WEIGHT
        p.prediction_time_span = self.PREDICTION_TIME_SPAN
        p.desired_velocity = self.DESIRED_VELOCITY
        p.discount_factor = self.DISCOUNT_FACTOR
        p.goal_reward = self.GOAL_REWARD
        p.reward_vector_size = self.REWARD_VECTOR_SIZE
        p.horizon = self.HORIZON
        p.use_rule_reward_for_ego_only = 1 if self.USE_RULE_REWARD_FOR_EGO_ONLY else 0
        p.remove_agents_out_of_map = 1 if self.REMOVE_AGENTS_OUT_OF_MAP else 0
        p.wheel_base = self.WHEEL_BASE
        p.integration_time_delta = self.INTEGRATION_TIME_DELTA
        p.sd_reaction_time = self.SAFE_DISTANCE_REACTION_TIME
        p.sd_max_decel_rear = self.SAFE_DISTANCE_MAX_DECEL_REAR
        p.sd_max_decel_front = self.SAFE_DISTANCE_MAX_DECEL_FRONT
        p.near_distance = self.NEAR_DISTANCE
        return p


@dataclass
class AgentModel:
    """Shape, goal and motion primitives (a, delta) of one agent.  Agents outside the MCTS
    agent list only ever use primitive 0 (bark BehaviorConstantAcceleration)."""
    front: float = 3.0
    rear: float = 1.0
    half_width: float = 0.9
    primitives: List[tuple] = field(default_factory=lambda: [(0.0, 0.0)])
    goal: Optional[tuple] = None  # (xmin, xmax, ymin, ymax[, theta_lo, theta_hi])

    def to_abi(self):
        a = _abi.CwAgent()
        a.front, a.rear, a.half_width = self.front, self.rear, self.half_width
        a.has_goal = 0 if self.goal is None else 1
        g = [0.0, 0.0, 0.0, 0.0, -BIG, BIG]
        if self.goal is not None:
            g[:len(self.goal)] = self.goal
        for k in range(6):
            a.goal[k] = g[k]
        a.n_prims = len(self.primitives)
        for k, (acc, delta) in enumerate(self.primitives):
            a.prim_acc[k] = acc
            a.prim_delta[k] = delta
        return a


@dataclass
class Lane:
    points: np.ndarray            # [P][2] centre line
    half_width: float = 1.75
    successor: int = -1
    s_point: float = BIG          # arc length of the merge point on this lane
    rightmost: bool = False

    def to_abi(self):
        l = _abi.CwLane()
        pts = np.asarray(self.points, np.float64)
        l.n_pts = pts.shape[0]
        l.successor = self.successor
        l.flags = 1 if self.rightmost else 0
        l.half_width = self.half_width
        l.s_point = self.s_point
        for k in range(pts.shape[0]):
            l.pts[k][0] = pts[k, 0]
            l.pts[k][1] = pts[k, 1]
        return l


@dataclass
class Scenario:
    params: RulesMctsStateParameters
    agents: List[AgentModel]
    n_mcts_agents: int
    lanes: List[Lane]
    road: np.ndarray              # [E][2] road polygon
    rules: List[RuleMonitor]
    # initial world state: [A][4] (x, y, theta, v)
    initial_state: Optional[np.ndarray] = None
    # per MCTS agent the indices of the rules that apply to it (None: every rule, AddCommonRules);
    # AddAgentRules of the reference: rules listed for some agents only.  The slots of the other
    # rules start (and stay) in the rule's `disabled_state` -- see initial_q().
    agent_rules: Optional[List[Optional[List[int]]]] = None

    def to_abi(self):
        s = _abi.CwScenario()
        s.params = self.params.to_abi()
        s.n_agents = len(self.agents)
        s.n_mcts_agents = self.n_mcts_agents
        s.n_lanes = len(self.lanes)
        road = np.asarray(self.road, np.float64)
        s.n_road = road.shape[0]
        s.n_rules = len(self.rules)
        for i, a in enumerate(self.agents):
            s.agents[i] = a.to_abi()
        for i, l in enumerate(self.lanes):
            s.lanes[i] = l.to_abi()
        for k in range(road.shape[0]):
            s.road[k][0] = road[k, 0]
            s.road[k][1] = road[k, 1]
        for k, r in enumerate(self.rules):
            s.rules[k] = r.to_abi(CW_LABELS)
        return s


def _is_torch(x):
    return type(x).__module__.startswith("torch")


def _ptr(x):
    if x is None:
        return None
    if _is_torch(x):
        return x.data_ptr()
    return x.ctypes.data


class CwBatch:
    """n states, structure-of-arrays (brm_cw_states in include/brm.h); numpy or torch CUDA."""

    FIELDS = ("scenario", "agents", "flags", "horizon", "terminal", "q", "viol")

    def __init__(self, scenario, agents, flags, horizon, terminal, q, viol=None):
        self.scenario, self.agents, self.flags = scenario, agents, flags
        self.horizon, self.terminal, self.q, self.viol = horizon, terminal, q, viol
        self.device = _is_torch(agents) and agents.is_cuda

    @property
    def n(self):
        return int(self.agents.shape[1])

    @staticmethod
    def zeros(A, M, R, n, track_violations=True):
        return CwBatch(np.zeros(n, np.int32), np.zeros((A, n, 4), np.float64), np.zeros((A, n), np.uint8),
                       np.zeros(n, np.int32), np.zeros(n, np.uint8), np.zeros((M, R, n), np.uint8),
                       np.zeros((M, R, n), np.uint32) if track_violations else None)

    def to_abi(self):
        s = _abi.CwStates()
        s.n = self.n
        s.mem = _abi.MEM_DEVICE if self.device else _abi.MEM_HOST
        for f in self.FIELDS:
            setattr(s, f, _ptr(getattr(self, f)))
        return s

    def cuda(self, device=0):
        import torch
        dev = torch.device("cuda", device)
        t = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        viol = None if self.viol is None else torch.from_numpy(self.viol.view(np.int32)).to(dev)
        return CwBatch(t(self.scenario), t(self.agents), t(self.flags), t(self.horizon), t(self.terminal),
                       t(self.q), viol)

    @staticmethod
    def concat(batches):
        """Host batches only: one batch holding the states of all `batches`, in order."""
        cat = lambda f, ax: np.ascontiguousarray(np.concatenate([getattr(b, f) for b in batches], axis=ax))
        viol = None if any(b.viol is None for b in batches) else cat("viol", 2)
        return CwBatch(cat("scenario", 0), cat("agents", 1), cat("flags", 1), cat("horizon", 0), cat("terminal", 0),
                       cat("q", 2), viol)

    def select(self, idx):
        idx = np.asarray(idx)
        v = None if self.viol is None else np.ascontiguousarray(self.viol[:, :, idx])
        return CwBatch(np.ascontiguousarray(self.scenario[idx]), np.ascontiguousarray(self.agents[:, idx]),
                       np.ascontiguousarray(self.flags[:, idx]), np.ascontiguousarray(self.horizon[idx]),
                       np.ascontiguousarray(self.terminal[idx]), np.ascontiguousarray(self.q[:, :, idx]), v)


class RulesMctsEngine:
    """Batched MvmctsState surface + rollouts for a set of scenarios on one GPU."""

    def __init__(self, scenarios, device=0, cuda_stream=None, engine=None):
        self.scenarios = list(scenarios) if isinstance(scenarios, (list, tuple)) else [scenarios]
        self.engine = engine or _abi.Engine(device, cuda_stream)
        self.lib = self.engine.lib
        s0 = self.scenarios[0]
        self.A, self.M, self.V = len(s0.agents), s0.n_mcts_agents, s0.params.REWARD_VECTOR_SIZE
        arr = (_abi.CwScenario * len(self.scenarios))()
        for i, s in enumerate(self.scenarios):
            arr[i] = s.to_abi()
        _abi.check(self.lib.brm_cw_configure(self.engine.handle, arr, len(self.scenarios)))
        self.R = int(self.lib.brm_cw_rules_per_agent(self.engine.handle))

    def initial_q(self, scenario=0):
        """[M][R] initial automaton states; rules that do not apply to an agent (Scenario.agent_rules)
        sit in their absorbing `disabled_state`."""
        scn = self.scenarios[scenario]
        q = np.zeros((self.M, self.R), np.uint8)
        s = 0
        for k, r in enumerate(scn.rules):
            reps = self.A - 1 if r.IsAgentSpecific() else 1
            q[:, s:s + reps] = r.table.init
            for a, lst in enumerate(scn.agent_rules or []):
                if lst is not None and k not in lst:
                    q[a, s:s + reps] = r.disabled_state
            s += reps
        return q

    def make_batch(self, states, scenario=None, flags=None, horizon=None, q=None, track_violations=True):
        """states [n][A][4] (x, y, theta, v); fills `terminal` with CheckTerminal on the GPU."""
        states = np.asarray(states, np.float64)
        if states.ndim == 2:
            states = states[None]
        n = states.shape[0]
        b = CwBatch.zeros(self.A, self.M, self.R, n, track_violations)
        b.agents[:] = states.transpose(1, 0, 2)
        b.scenario[:] = 0 if scenario is None else scenario
        b.flags[:] = (PRESENT | VALID) if flags is None else np.asarray(flags, np.uint8).reshape(n, self.A).T
        sc = np.atleast_1d(b.scenario)
        if horizon is None:
            b.horizon[:] = [self.scenarios[int(k)].params.HORIZON for k in sc]
        else:
            b.horizon[:] = horizon
        if q is None:
            for k in np.unique(sc):
                b.q[:, :, sc == k] = self.initial_q(int(k))[:, :, None]
        else:
            q = np.asarray(q, np.uint8)
            b.q[:] = q[:, :, None] if q.ndim == 2 else q.transpose(1, 2, 0)
        self.check_terminal(b)
        return b

    def initial_batch(self, scenario_ids=None):
        """One root state per scenario (its `initial_state`)."""
        ids = list(range(len(self.scenarios))) if scenario_ids is None else list(scenario_ids)
        states = np.stack([self.scenarios[k].initial_state for k in ids])
        b = None
        # CheckTerminal needs scenario-homogeneous batches
        parts = [self.make_batch(states[i:i + 1], scenario=ids[i]) for i in range(len(ids))]
        if len(parts) == 1:
            return parts[0]
        cat = lambda f, ax: np.concatenate([getattr(p, f) for p in parts], axis=ax)
        b = CwBatch(cat("scenario", 0), cat("agents", 1), cat("flags", 1), cat("horizon", 0),
                    cat("terminal", 0), cat("q", 2), cat("viol", 2))
        return b

    # -- StateInterface surface, batched ------------------------------------------------
    def check_terminal(self, batch):
        s = batch.to_abi()
        _abi.check(self.lib.brm_cw_check_terminal(self.engine.handle, C.byref(s)))
        return batch.terminal

    def is_terminal(self, batch):
        """The terminal flags the states carry (set by Execute / check_terminal)."""
        return np.asarray(batch.terminal).astype(bool)

    def execute(self, batch, actions):
        """MvmctsState::Execute; actions [M][n] uint8 -> (next_batch, rewards [M][V][n] float64)."""
        n = batch.n
        if batch.device:
            import torch
            out = CwBatch(*[None if getattr(batch, f) is None else torch.empty_like(getattr(batch, f))
                            for f in CwBatch.FIELDS])
            rewards = torch.zeros((self.M, self.V, n), dtype=torch.float64, device=batch.agents.device)
        else:
            actions = np.ascontiguousarray(actions, np.uint8)
            out = CwBatch.zeros(self.A, self.M, self.R, n, batch.viol is not None)
            rewards = np.zeros((self.M, self.V, n), np.float64)
        sin, sout = batch.to_abi(), out.to_abi()
        _abi.check(self.lib.brm_cw_execute(self.engine.handle, C.byref(sin), _ptr(actions), C.byref(sout),
                                           _ptr(rewards)))
        return out, rewards

    def evaluate_rules(self, batch):
        """MvmctsState::EvaluateRules(): one automaton transition on the labels of the states as
        they are -> (batch with the new rule states, rule penalties [M][V][n] float64)."""
        n = batch.n
        if batch.device:
            import torch
            out = CwBatch(*[None if getattr(batch, f) is None else torch.empty_like(getattr(batch, f))
                            for f in CwBatch.FIELDS])
            rewards = torch.zeros((self.M, self.V, n), dtype=torch.float64, device=batch.agents.device)
        else:
            out = CwBatch.zeros(self.A, self.M, self.R, n, batch.viol is not None)
            rewards = np.zeros((self.M, self.V, n), np.float64)
        sin, sout = batch.to_abi(), out.to_abi()
        _abi.check(self.lib.brm_cw_evaluate_rules(self.engine.handle, C.byref(sin), C.byref(sout), _ptr(rewards)))
        return out, rewards

    def get_num_actions(self, batch):
        out = np.zeros((self.M, batch.n), np.uint8)
        s = batch.to_abi()
        _abi.check(self.lib.brm_cw_get_num_actions(self.engine.handle, C.byref(s), _ptr(out)))
        return out

    def terminal_reward(self, batch):
        out = np.zeros((self.M, self.V, batch.n), np.float64)
        s = batch.to_abi()
        _abi.check(self.lib.brm_cw_terminal_reward(self.engine.handle, C.byref(s), _ptr(out)))
        return out

    # -- rollouts -----------------------------------------------------------------------
    def rollout(self, leaves, rollouts_per_leaf, seed=1000, rollout_base=0, max_steps=30,
                discount_factor=None, first_discount_exp=1, detail=False, out=None, coop_factor=0.0,
                first_action=None, stats=None):
        """Random rollouts from every leaf.  With `first_action` [L][M] and `stats`
        [M][n_actions][1+V] the per-leaf sums are also folded into the root statistics in the same
        call (brm_cw_rollout_root_stats)."""
        gamma = self.scenarios[0].params.DISCOUNT_FACTOR if discount_factor is None else discount_factor
        rp = _abi.RolloutParams(seed, rollout_base, rollouts_per_leaf, max_steps, gamma, first_discount_exp, 0,
                                coop_factor)
        L = leaves.n
        N = L * rollouts_per_leaf
        A, M, V, R = self.A, self.M, self.V, self.R
        if out is None:
            assert not leaves.device, "device batches need preallocated outputs"
            out = dict(returns_sum=np.zeros((L, M, V), np.float64), viol_sum=np.zeros((L, M, R), np.uint64),
                       agent_steps=np.zeros(L, np.uint64))
            if detail:
                out.update(ro_returns=np.zeros((N, M, V), np.float64), ro_viol=np.zeros((N, M, R), np.uint8),
                           ro_q=np.zeros((N, M, R), np.uint8), ro_agents=np.zeros((N, A, 4), np.float64),
                           ro_flags=np.zeros((N, A), np.uint8), ro_steps=np.zeros(N, np.int32))
        o = _abi.CwRolloutOut()
        for k, _ in _abi.CwRolloutOut._fields_:
            setattr(o, k, _ptr(out.get(k)))
        s = leaves.to_abi()
        if stats is not None:
            _abi.check(self.lib.brm_cw_rollout_root_stats(self.engine.handle, C.byref(s), C.byref(rp), C.byref(o),
                                                          _ptr(first_action), int(stats.shape[1]), _ptr(stats)))
        else:
            _abi.check(self.lib.brm_cw_rollout(self.engine.handle, C.byref(s), C.byref(rp), C.byref(o)))
        return out

    # -- replay ---------------------------------------------------------------------------
    def replay(self, batch0, actions):
        """actions [B][T][M] uint8 -> dict of per-step traces (brm_cw_trace)."""
        actions = np.ascontiguousarray(actions, np.uint8)
        B, T, M = actions.shape
        assert B == batch0.n and M == self.M and not batch0.device
        A, V, R = self.A, self.V, self.R
        tr = dict(agents_out=np.zeros((B, T, A, 4), np.float64), flags_out=np.zeros((B, T, A), np.uint8),
                  terminal_out=np.zeros((B, T), np.uint8), q_out=np.zeros((B, T, M, R), np.uint8),
                  viol_out=np.zeros((B, T, M, R), np.uint32), rewards_out=np.zeros((B, T, M, V), np.float64),
                  labels_out=np.zeros((B, T, M, 4), np.uint32), frenet_out=np.zeros((B, T, A, 3), np.float64),
                  term_reward_out=np.zeros((B, M, V), np.float64), steps_out=np.zeros(B, np.int32))
        t = _abi.CwTrace()
        for k, _ in _abi.CwTrace._fields_:
            setattr(t, k, _ptr(tr[k]))
        s = batch0.to_abi()
        _abi.check(self.lib.brm_cw_replay(self.engine.handle, C.byref(s), _ptr(actions), T, C.byref(t)))
        return tr

    def root_stats_reduce(self, leaf_first_action, returns_sum, rollouts_per_leaf, stats, n_actions):
        dev = _is_torch(stats)
        L = int(leaf_first_action.shape[0])
        _abi.check(self.lib.brm_root_stats_reduce(
            self.engine.handle, _abi.MEM_DEVICE if dev else _abi.MEM_HOST, L, self.M, n_actions, self.V,
            _ptr(leaf_first_action), _ptr(returns_sum), rollouts_per_leaf, _ptr(stats)))
        return stats


class MvmctsState:
    """Single-state view with the reference's member names (src/rules_mcts_state.hpp:51-70)."""

    ego_agent_idx = 0

    def __init__(self, engine: RulesMctsEngine, batch: CwBatch = None, scenario=0):
        self.engine = engine
        self.batch = batch if batch is not None else engine.initial_batch([scenario])

    def Clone(self):
        b = self.batch
        return MvmctsState(self.engine, CwBatch(*[None if getattr(b, f) is None else getattr(b, f).copy()
                                                  for f in CwBatch.FIELDS]))

    def Execute(self, joint_action, rewards=None):
        """With a `rewards` list (the C++ out-parameter) returns the successor; without one
        returns (successor, rewards) like the reference's Python binding
        (python/python_planner_rules_mcts.cpp:113-120)."""
        acts = np.asarray(joint_action, np.uint8).reshape(self.engine.M, 1)
        nb, rew = self.engine.execute(self.batch, acts)  # raises BrmError(ERR_TERMINAL) on a terminal state
        out = [rew[a, :, 0].astype(np.float64) for a in range(self.engine.M)]
        if rewards is not None:
            rewards[:] = out
            return MvmctsState(self.engine, nb)
        return MvmctsState(self.engine, nb), out

    def EvaluateRules(self):
        """:186-197: transit the rule automata on this state (mutates it), returns the penalties."""
        nb, rew = self.engine.evaluate_rules(self.batch)
        self.batch = nb
        return [rew[a, :, 0].astype(np.float64) for a in range(self.engine.M)]

    @property
    def observed_world(self):
        """What is left of bark's ObservedWorld here: agent states [A][4] and flags [A]."""
        return self.GetAgentStates(), self.GetAgentFlags()

    def GetNumActions(self, agent_idx):
        return int(self.engine.get_num_actions(self.batch)[agent_idx, 0])

    def IsTerminal(self):
        return bool(self.batch.terminal[0])

    def GetAgentIdx(self):
        return list(range(self.engine.M))

    def PrintState(self):
        return ""  # src/rules_mcts_state.cpp:146

    def GetTerminalReward(self):
        tr = self.engine.terminal_reward(self.batch)
        return [tr[a, :, 0].astype(np.float64) for a in range(self.engine.M)]

    def GetMultiAgentRuleState(self):
        return self.batch.q[:, :, 0].copy()

    def GetAgentStates(self):
        return self.batch.agents[:, 0, :].copy()

    def GetAgentFlags(self):
        return self.batch.flags[:, 0].copy()
