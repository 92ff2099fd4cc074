"""Host mirror of the reference's GridWorld state plugin on top of the C ABI.

`GridWorldState` keeps the member set of mvmcts::StateInterface as implemented by
/root/reference/gridworld/grid_world_state.{hpp,cpp} (Clone, Execute, GetNumActions,
IsTerminal, GetAgentIdx, PrintState, GetTerminalReward, GetAgentStates, ResetDepth,
EgoGoalReached, GetEgoPos); `GridWorldEngine` is the batched surface a host MCTS drives
(Execute / rollouts for many states at once).  All arithmetic runs in libbrm.so on the
GPU; nothing here computes a step on the host.
"""
import ctypes as C
from dataclasses import dataclass, field
from typing import List

import numpy as np

from . import _abi
from .ltl import RuleMonitor

# label names of gridworld/base_test_env.cpp:104-117 -> C-ABI label kinds
GW_LABELS = {
    "collision": _abi.GW_LABEL_COLLISION,
    "merged_e": _abi.GW_LABEL_MERGED_E,
    "merged_x": _abi.GW_LABEL_MERGED_X,
    "in_direct_front_x": _abi.GW_LABEL_IN_DIRECT_FRONT_X,
}

# gridworld/base_test_env.cpp:16-19
ZIP_FORMULA = ("(in_direct_front_x#0 & !merged_e & (in_direct_front_x#0 | merged_x#0) U "
               "merged_e) -> G(merged_e & merged_x#0 -> !in_direct_front_x#0)")


@dataclass
class GridWorldStateParameter:
    """gridworld/grid_world_state_parameter.h:15-29; defaults of
    BaseTestEnv::MakeGridWorldStateParameters (gridworld/base_test_env.cpp:79-97)."""
    num_other_agents: int = 2
    state_x_length: int = 1000
    ego_goal_reached_position: int = 1000
    terminal_depth_: int = 8
    merging_point: int = 8
    speed_deviation_weight: float = 5.0
    acceleration_weight: float = 1.5
    potential_weight: float = 0.0
    reward_vec_size: int = 3
    action_map: List[int] = field(default_factory=lambda: [1, 0, -1])  # FORWARD, WAIT, BACKWARD

    def to_abi(self):
        p = _abi.GwParams()
        p.n_agents = self.num_other_agents + 1
        p.state_x_length = self.state_x_length
        p.ego_goal_reached_position = self.ego_goal_reached_position
        p.terminal_depth = self.terminal_depth_
        p.merging_point = self.merging_point
        p.reward_vec_size = self.reward_vec_size
        p.n_actions = len(self.action_map)
        for k, d in enumerate(self.action_map):
            p.action_map[k] = d
        p.speed_deviation_weight = self.speed_deviation_weight
        p.acceleration_weight = self.acceleration_weight
        p.potential_weight = self.potential_weight
        return p


def make_default_rules(on_violation="reset"):
    """BaseTestEnv::MakeAutomata (gridworld/base_test_env.cpp:98-109)."""
    return [RuleMonitor.MakeRule("G !collision", -1.0, 0, on_violation),
            RuleMonitor.MakeRule(ZIP_FORMULA, -1.0, 1, on_violation)]


def _is_torch(x):
    return type(x).__module__.startswith("torch")


def _ptr(x):
    if x is None:
        return None
    if _is_torch(x):
        return x.data_ptr()
    return x.ctypes.data


class GwBatch:
    """n GridWorld states, structure-of-arrays (see brm_gw_states in include/brm.h).
    Arrays are numpy (host) or torch CUDA tensors (device)."""

    def __init__(self, agents, term, depth, q, viol=None):
        self.agents, self.term, self.depth, self.q, self.viol = agents, term, depth, q, viol
        self.device = _is_torch(agents) and agents.is_cuda

    @property
    def n(self):
        return int(self.agents.shape[1])

    @staticmethod
    def zeros(A, R, n, track_violations=True):
        return GwBatch(np.zeros((A, n, 4), np.int32), np.zeros(n, np.uint8), np.zeros(n, np.int32),
                       np.zeros((A, R, n), np.uint8),
                       np.zeros((A, R, n), np.uint32) if track_violations else None)

    def to_abi(self):
        s = _abi.GwStates()
        s.n = self.n
        s.mem = _abi.MEM_DEVICE if self.device else _abi.MEM_HOST
        s.agents, s.term, s.depth, s.q = _ptr(self.agents), _ptr(self.term), _ptr(self.depth), _ptr(self.q)
        s.viol = _ptr(self.viol)
        return s

    def cuda(self, device=0):
        import torch
        dev = torch.device("cuda", device)
        t = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        # torch has no uint32 arithmetic but can hold the bytes: view counters as int32
        viol = None if self.viol is None else torch.from_numpy(self.viol.view(np.int32)).to(dev)
        return GwBatch(t(self.agents), t(self.term), t(self.depth), t(self.q), viol)

    @staticmethod
    def concat(batches):
        """Host batches only: one batch holding the states of all `batches`, in order."""
        cat = lambda f, ax: np.ascontiguousarray(np.concatenate([getattr(b, f) for b in batches], axis=ax))
        viol = None if any(b.viol is None for b in batches) else cat("viol", 2)
        return GwBatch(cat("agents", 1), cat("term", 0), cat("depth", 0), cat("q", 2), viol)

    def select(self, idx):
        """Host batches only: states idx (array of indices)."""
        v = None if self.viol is None else np.ascontiguousarray(self.viol[:, :, idx])
        return GwBatch(np.ascontiguousarray(self.agents[:, idx]), np.ascontiguousarray(self.term[idx]),
                       np.ascontiguousarray(self.depth[idx]), np.ascontiguousarray(self.q[:, :, idx]), v)


class GridWorldEngine:
    """Batched GridWorld StateInterface + rollout surface on one GPU."""

    def __init__(self, params=None, rules=None, device=0, cuda_stream=None, engine=None):
        self.params = params or GridWorldStateParameter()
        self.rules = list(rules) if rules is not None else make_default_rules()
        self.engine = engine or _abi.Engine(device, cuda_stream)
        self.lib = self.engine.lib
        self.A = self.params.num_other_agents + 1
        self.M = self.A  # every GridWorld agent is an MCTS agent
        self.V = self.params.reward_vec_size
        abi_rules = (_abi.Rule * max(1, len(self.rules)))()
        for k, r in enumerate(self.rules):
            abi_rules[k] = r.to_abi(GW_LABELS)
        self._abi_rules = abi_rules
        p = self.params.to_abi()
        _abi.check(self.lib.brm_gw_configure(self.engine.handle, C.byref(p), abi_rules, len(self.rules)))
        self.R = int(self.lib.brm_gw_rules_per_agent(self.engine.handle))

    # -- state construction -------------------------------------------------------
    def initial_q(self):
        """[A][R] initial automaton states in slot order (rules in order, agent-specific
        rules once per other agent ascending; gridworld/base_test_env.cpp:122-140)."""
        q = np.zeros((self.A, self.R), np.uint8)
        for a in range(self.A):
            s = 0
            for r in self.rules:
                reps = self.A - 1 if r.IsAgentSpecific() else 1
                q[a, s:s + reps] = r.table.init
                s += reps
        return q

    def make_batch(self, agents, term=None, depth=None, q=None, track_violations=True):
        """agents: [n][A][4] or [A][4] (x_pos, last_action, lane, init_lane)."""
        agents = np.asarray(agents, np.int32)
        if agents.ndim == 2:
            agents = agents[None]
        n = agents.shape[0]
        b = GwBatch.zeros(self.A, self.R, n, track_violations)
        b.agents[:] = agents.transpose(1, 0, 2)
        if term is not None:
            t = np.asarray(term, np.uint8).reshape(n, self.A)
            b.term[:] = (t.astype(np.uint32) << np.arange(self.A, dtype=np.uint32)).sum(1).astype(np.uint8)
        if depth is not None:
            b.depth[:] = depth
        q0 = self.initial_q() if q is None else np.asarray(q, np.uint8)
        b.q[:] = q0[:, :, None] if q0.ndim == 2 else q0.transpose(1, 2, 0)
        return b

    def base_test_env_batch(self, n=1):
        """Initial state of BaseTestEnv (gridworld/base_test_env.cpp:29-43), n copies."""
        mp = self.params.merging_point
        ag = np.array([[mp - 6, 1, 1, 1], [mp - 2, 1, 1, 1], [mp - 8, 1, 3, 3]], np.int32)
        return self.make_batch(np.broadcast_to(ag, (n, 3, 4)))

    # -- StateInterface surface, batched --------------------------------------------
    def execute(self, batch, actions):
        """GridWorldState::Execute for every state; actions [A][n] uint8.
        Returns (next_batch, rewards [A][V][n] float64)."""
        n = batch.n
        if batch.device:
            import torch
            out = GwBatch(torch.empty_like(batch.agents), torch.empty_like(batch.term),
                          torch.empty_like(batch.depth), torch.empty_like(batch.q),
                          None if batch.viol is None else torch.zeros_like(batch.viol))
            rewards = torch.empty((self.A, self.V, n), dtype=torch.float64, device=batch.agents.device)
        else:
            actions = np.ascontiguousarray(actions, np.uint8)
            out = GwBatch.zeros(self.A, self.R, n, batch.viol is not None)
            rewards = np.zeros((self.A, self.V, n), np.float64)
        sin, sout = batch.to_abi(), out.to_abi()
        _abi.check(self.lib.brm_gw_execute(self.engine.handle, C.byref(sin), _ptr(actions),
                                           C.byref(sout), _ptr(rewards)))
        return out, rewards

    def get_num_actions(self, batch):
        out = np.zeros((self.A, batch.n), np.uint8)
        s = batch.to_abi()
        _abi.check(self.lib.brm_gw_get_num_actions(self.engine.handle, C.byref(s), _ptr(out)))
        return out

    def is_terminal(self, batch):
        out = np.zeros(batch.n, np.uint8)
        s = batch.to_abi()
        _abi.check(self.lib.brm_gw_is_terminal(self.engine.handle, C.byref(s), _ptr(out)))
        return out.astype(bool)

    def terminal_reward(self, batch):
        out = np.zeros((self.A, self.V, batch.n), np.float64)
        s = batch.to_abi()
        _abi.check(self.lib.brm_gw_terminal_reward(self.engine.handle, C.byref(s), _ptr(out)))
        return out

    # -- rollouts (mvmcts::RandomHeuristic) ----------------------------------------------
    def rollout(self, leaves, rollouts_per_leaf, seed=1000, rollout_base=0, max_steps=30,
                discount_factor=0.95, first_discount_exp=1, detail=False, out=None, coop_factor=0.0,
                first_action=None, stats=None):
        """Random rollouts from every leaf state.  Host batches return numpy arrays; for a
        device batch pass `out` = dict of preallocated CUDA tensors (returns_sum, viol_sum,
        agent_steps) and nothing is copied.  coop_factor: mvmcts COOP_FACTOR, returns_sum of
        agent i = own_i + coop_factor * sum of the others'.  With `first_action` [L][A] and
        `stats` [A][n_actions][1+V] the per-leaf sums are also folded into the root statistics in
        the same call (brm_gw_rollout_root_stats)."""
        rp = _abi.RolloutParams(seed, rollout_base, rollouts_per_leaf, max_steps, discount_factor,
                                first_discount_exp, 0, coop_factor)
        L = leaves.n
        N = L * rollouts_per_leaf
        A, V, R = self.A, self.V, self.R
        if out is None:
            assert not leaves.device, "device batches need preallocated outputs"
            out = dict(returns_sum=np.zeros((L, A, V), np.float64),
                       viol_sum=np.zeros((L, A, R), np.uint64),
                       agent_steps=np.zeros(L, np.uint64))
            if detail:
                out.update(ro_returns=np.zeros((N, A, V), np.float64),
                           ro_viol=np.zeros((N, A, R), np.uint8), ro_q=np.zeros((N, A, R), np.uint8),
                           ro_agents=np.zeros((N, A, 4), np.int32), ro_steps=np.zeros(N, np.int32))
        o = _abi.GwRolloutOut()
        for k, _ in _abi.GwRolloutOut._fields_:
            setattr(o, k, _ptr(out.get(k)))
        s = leaves.to_abi()
        if stats is not None:
            _abi.check(self.lib.brm_gw_rollout_root_stats(self.engine.handle, C.byref(s), C.byref(rp), C.byref(o),
                                                          _ptr(first_action), int(stats.shape[1]), _ptr(stats)))
        else:
            _abi.check(self.lib.brm_gw_rollout(self.engine.handle, C.byref(s), C.byref(rp), C.byref(o)))
        return out

    # -- replay (parity instrument) ---------------------------------------------------------
    def replay(self, batch0, actions):
        """actions [B][T][A] uint8 -> dict of per-step traces (see brm_gw_trace)."""
        actions = np.ascontiguousarray(actions, np.uint8)
        B, T, A = actions.shape
        assert B == batch0.n and A == self.A and not batch0.device
        V, R = self.V, self.R
        tr = dict(agents_out=np.zeros((B, T, A, 4), np.int32), term_out=np.zeros((B, T), np.uint8),
                  q_out=np.zeros((B, T, A, R), np.uint8), viol_out=np.zeros((B, T, A, R), np.uint32),
                  rewards_out=np.zeros((B, T, A, V), np.float64),
                  labels_out=np.zeros((B, T, A, 3), np.uint32),
                  term_reward_out=np.zeros((B, A, V), np.float64), steps_out=np.zeros(B, np.int32))
        t = _abi.GwTrace()
        for k, _ in _abi.GwTrace._fields_:
            setattr(t, k, _ptr(tr[k]))
        s = batch0.to_abi()
        _abi.check(self.lib.brm_gw_replay(self.engine.handle, C.byref(s), _ptr(actions), T, C.byref(t)))
        return tr

    # -- root statistics ------------------------------------------------------------------
    def root_stats_reduce(self, leaf_first_action, returns_sum, rollouts_per_leaf, stats, n_actions=None):
        """Accumulate per-leaf rollout sums into root statistics [A][n_actions][1+V]."""
        n_actions = n_actions or len(self.params.action_map)
        dev = _is_torch(stats)
        L = int(leaf_first_action.shape[0])
        _abi.check(self.lib.brm_root_stats_reduce(
            self.engine.handle, _abi.MEM_DEVICE if dev else _abi.MEM_HOST, L, self.A, n_actions,
            self.V, _ptr(leaf_first_action), _ptr(returns_sum), rollouts_per_leaf, _ptr(stats)))
        return stats


class GridWorldState:
    """Single-state view with the reference's StateInterface member names
    (gridworld/grid_world_state.hpp:28-71); each call runs a batch of one on the GPU."""

    ego_agent_idx = 0

    def __init__(self, engine: GridWorldEngine, batch: GwBatch = None):
        self.engine = engine
        self.batch = batch if batch is not None else engine.base_test_env_batch(1)

    def Clone(self):
        b = self.batch
        return GridWorldState(self.engine, GwBatch(b.agents.copy(), b.term.copy(), b.depth.copy(),
                                                   b.q.copy(), None if b.viol is None else b.viol.copy()))

    def Execute(self, joint_action, rewards=None):
        """Returns the successor state; `rewards` (a list) is resized and filled with one
        reward vector per agent, as the reference's out-parameter."""
        acts = np.asarray(joint_action, np.uint8).reshape(self.engine.A, 1)
        nb, rew = self.engine.execute(self.batch, acts)
        if rewards is not None:
            rewards[:] = [rew[a, :, 0].copy() for a in range(self.engine.A)]
        return GridWorldState(self.engine, nb)

    def GetTerminalReward(self):
        tr = self.engine.terminal_reward(self.batch)
        return [tr[a, :, 0].copy() for a in range(self.engine.A)]

    def GetNumActions(self, agent_idx):
        return int(self.engine.get_num_actions(self.batch)[agent_idx, 0])

    def IsTerminal(self):
        return bool(self.engine.is_terminal(self.batch)[0])

    def GetAgentIdx(self):
        return list(range(self.engine.A))

    def GetAgentStates(self):
        return self.batch.agents[:, 0, :].copy()

    def GetEgoPos(self):
        return int(self.batch.agents[0, 0, 0])

    def EgoGoalReached(self):
        return self.GetEgoPos() >= self.engine.params.ego_goal_reached_position

    def ResetDepth(self):
        self.batch.depth[:] = 0

    def GetRuleStates(self):
        return self.batch.q[:, :, 0].copy()

    def PrintState(self):
        x = self.batch.agents[:, 0, 0]
        s = "Ego: x=%d" % x[0]
        for i in range(1, len(x)):
            s += ", Ag%d: x=%d" % (i, x[i])
        return s + "\n"
