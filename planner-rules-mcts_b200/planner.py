"""Host planner shell on top of the device rollout engine (SURVEY.md 8f, row N1 -- the caller
of the hot path).

`BehaviorRulesMcts` keeps the outer surface of the reference's planner plugin
(/root/reference/src/behavior_rules_mcts.hpp:60-119: constructed from parameters, rules and
prediction settings; `Plan()` searches from the observed world and returns the best action of
the ego agent) but the search itself is a *root-parallel flat Monte-Carlo* search: every
joint root action (or a sample of them) becomes a leaf, the leaves are evaluated by batched
device rollouts (mvmcts::RandomHeuristic, fused in K1/K3), the per-agent root statistics are
folded on the device (K5) and, with several ranks, merged by one allreduce.  The reference's
mvmcts tree with thresholded lexicographic UCT statistics is an un-vendored dependency; only
its root-level pick (`parallel.best_actions`) is mirrored.  Behavioural pins reproduced in
tests/test_planner_gpu.py: test/single_agent_test.cpp:229-270 (accelerate when the road is
free, brake behind a slower agent) and the closed loop of :272-306 (goal reached without a
collision).
"""
import itertools
from dataclasses import dataclass

import numpy as np

from . import parallel
from .continuous import RulesMctsEngine, Scenario


@dataclass
class PlannerParameters:
    """BehaviorRulesMcts::* parameters (src/behavior_rules_mcts.hpp:134-153, src/util.cpp:23-89)."""
    MaxNumIterations: int = 2000          # rollouts per decision here (one rollout per iteration there)
    RandomSeed: int = 1000
    MaxNumIterationsRandomHeuristic: int = 30
    MaxJointActions: int = 243            # flat search: leaves per decision at most (sampled beyond that)
    Threshold: tuple = None               # thresholded lexicographic pick; None = plain lexicographic
    Search: str = "mcts"                  # "mcts": host tree (mcts.py); "flat": depth-1 root-parallel search
    RolloutsPerExpansion: int = 16        # mcts: device rollouts per expanded node
    UCTExplorationConstant: float = 1.42  # src/util.cpp:51-56
    Statistic: str = "uct"                # "uct" (BehaviorRulesMctsUct) | "greedy" (BehaviorRulesMctsGreedy)
    BatchSize: int = 1                    # mcts: descents per device round (1 = the reference's serial loop)


class BehaviorRulesMcts:
    def __init__(self, scenario: Scenario, params: PlannerParameters = None, device=0, engine=None):
        self.scenario = scenario
        self.params = params or PlannerParameters()
        self.engine = engine or RulesMctsEngine([scenario], device=device)
        self.decision = 0
        self.last_action = None
        self.last_stats = None

    def _root_joint_actions(self, n_act):
        M = self.engine.M
        total = int(np.prod(n_act))
        if total <= self.params.MaxJointActions:
            return np.array(list(itertools.product(*[range(k) for k in n_act])), np.uint8)
        rng = np.random.default_rng(self.params.RandomSeed + self.decision)
        ja = np.stack([rng.integers(0, n_act[a], size=self.params.MaxJointActions) for a in range(M)], axis=1)
        return ja.astype(np.uint8)

    def Plan(self, state_batch):
        """One decision from a batch of one state (CwBatch, n == 1).  Returns the ego action."""
        eng = self.engine
        if state_batch.terminal[0]:
            # LOG_IF(FATAL, mcts_state.IsTerminal()), src/behavior_rules_mcts.hpp:208-209
            raise RuntimeError("Current state is terminal! Aborting!")
        if self.params.Search == "mcts":
            from .mcts import Mcts, MctsParameters
            sp = self.scenario.params
            mp = MctsParameters(eng.V, discount_factor=sp.DISCOUNT_FACTOR,
                                exploration_constant=self.params.UCTExplorationConstant,
                                threshold=self.params.Threshold,
                                max_rollout_steps=self.params.MaxNumIterationsRandomHeuristic,
                                rollouts_per_expansion=self.params.RolloutsPerExpansion,
                                seed=self.params.RandomSeed + self.decision, statistic=self.params.Statistic,
                                batch_size=self.params.BatchSize)
            m = Mcts(eng, mp).Search(state_batch, self.params.MaxNumIterations)
            self.last_stats = m.GetRootStats()
            self.last_action = m.ReturnBestAction()[0]
            self.decision += 1
            return self.last_action
        n_act = eng.get_num_actions(state_batch)[:, 0]
        ja = self._root_joint_actions(n_act)
        L = ja.shape[0]
        roots = state_batch.select(np.zeros(L, np.int64))
        leaves, step_rewards = eng.execute(roots, ja.T.copy())
        per_leaf = max(1, self.params.MaxNumIterations // L)
        # root-parallel: every rank evaluates the root's joint actions with its own rollouts
        base = parallel.rollout_base(parallel.rank_and_world()[0], self.decision, L * per_leaf)
        gamma = self.scenario.params.DISCOUNT_FACTOR
        out = eng.rollout(leaves, per_leaf, seed=self.params.RandomSeed, rollout_base=base,
                          max_steps=self.params.MaxNumIterationsRandomHeuristic, discount_factor=gamma,
                          first_discount_exp=1)
        # value of a root action = r(root, a) + gamma * return of the leaf; RandomHeuristic already
        # weights its first step by gamma (first_discount_exp = 1)
        returns = out["returns_sum"] + per_leaf * step_rewards.transpose(2, 0, 1).astype(np.float64)
        n_actions = int(n_act.max())
        stats = np.zeros((eng.M, n_actions, 1 + eng.V))
        eng.root_stats_reduce(ja, np.ascontiguousarray(returns), per_leaf, stats, n_actions)
        parallel.merge_root_stats(stats)
        best = parallel.best_actions(stats, self.params.Threshold)
        self.last_action, self.last_stats = int(best[0]), stats
        self.decision += 1
        return int(best[0])

    def GetLastAction(self):
        return self.last_action

    def GetExpectedRewards(self):
        """Ego: root action -> mean return vector (GetEgoIntNode().GetExpectedRewards())."""
        q = parallel.expected_rewards(self.last_stats)
        return {a: q[0, a] for a in range(q.shape[1])}


def run_closed_loop(planner: BehaviorRulesMcts, max_steps=100, others_action=0):
    """World::Step loop of test/single_agent_test.cpp:286-299: plan, apply the ego action (the
    other MCTS agents, if any, apply `others_action`), repeat until the state is terminal.
    Returns the list of visited states [steps+1][A][4] and the final batch."""
    eng = planner.engine
    cur = eng.initial_batch([0])
    history = [cur.agents[:, 0, :].copy()]
    for _ in range(max_steps):
        if cur.terminal[0]:
            break
        a = planner.Plan(cur)
        act = np.full((eng.M, 1), others_action, np.uint8)
        act[0, 0] = a
        cur, _ = eng.execute(cur, act)
        # a new planning step starts with the full horizon again (Plan() builds a fresh
        # MvmctsState with state_params_.HORIZON, src/behavior_rules_mcts.hpp:205-207)
        if not cur.terminal[0] or cur.horizon[0] == 0:
            cur.horizon[:] = planner.scenario.params.HORIZON
            eng.check_terminal(cur)
        history.append(cur.agents[:, 0, :].copy())
    return np.array(history), cur


def write_state_history_tsv(path, history):
    """gridworld/state_file_writer.cpp:10-28 format: Timestep, then one x column per agent."""
    with open(path, "w") as f:
        f.write("Timestep" + "".join("\tAgent %d" % i for i in range(history.shape[1])) + "\n")
        for t, s in enumerate(history):
            f.write(str(t) + "".join("\t%g" % x for x in s[:, 0]) + "\n")
