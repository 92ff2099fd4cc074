"""TEST SCAFFOLDING (not part of the product package) -- GridWorld test environment and
closed-loop runner on the device engine (configs[0] of BASELINE.json).

Mirrors /root/reference/gridworld/base_test_env.{h,cpp} (`BaseTestEnv`: three agents before a
merging point, `G !collision` + zipper rule per agent, the mvmcts parameter set of
`MakeMctsParameters`), gridworld/grid_world_env.h (`CrossingTestEnv::Search`, `GetEgoQval`) and
gridworld/test_runner.cpp:17-52 (`TestRunner::RunTest`: search, execute the best joint action,
`ResetDepth`, accumulate rewards, stop at a terminal state or after `max_steps`).  Every
`Execute`, rollout and terminal reward runs in libbrm.so on the GPU; the search tree is the host
tree of mcts.py (the reference's ThresUCTStatistic is un-vendored: selection details are this
package's own, "parity unpinned"; what the reference's test pins are the three behaviours of
test/grid_world_threshold_test.cpp:40-71).
"""
from dataclasses import dataclass

import numpy as np

from planner_rules_mcts_b200.gridworld import (GridWorldEngine, GridWorldState, GridWorldStateParameter,
                                               make_default_rules)
from planner_rules_mcts_b200.mcts import Mcts, MctsParameters


def MakeMctsParameters(thres, rollouts_per_expansion=16, seed=1000, statistic="uct", batch_size=1):
    """BaseTestEnv::MakeMctsParameters (gridworld/base_test_env.cpp:45-77)."""
    thres = np.asarray(thres, np.float64)
    assert thres.size == 3
    return MctsParameters(
        3, discount_factor=0.95, exploration_constant=[0.8, 0.8, 0.8],
        lower_bound=[-1.0, -1.0, -100.0], upper_bound=[0.0, 0.0, 0.0], threshold=thres,
        max_rollout_steps=30, rollouts_per_expansion=rollouts_per_expansion, seed=seed,
        statistic=statistic, decay1=0.8, decay2=0.12, batch_size=batch_size)


class CrossingTestEnv:
    """BaseTestEnv + CrossingTestEnv<Stats>: owns the state, the MCTS and the histories."""

    def __init__(self, thres, device=0, engine=None, rollouts_per_expansion=16, seed=1000, statistic="uct",
                 batch_size=1):
        self.grid_world_state_parameter = GridWorldStateParameter()
        self.engine = engine or GridWorldEngine(self.grid_world_state_parameter, make_default_rules(),
                                                device=device)
        self.mcts_parameters = MakeMctsParameters(thres, rollouts_per_expansion, seed, statistic, batch_size)
        self.state = GridWorldState(self.engine)
        A, V = self.engine.A, self.engine.V
        self.rewards = np.zeros((A, V))
        self.state_history = []
        self.action_history = []
        self.jt = [0] * A                         # BaseTestEnv: jt_ starts as all-FORWARD
        self.mcts = None
        self._searches = 0

    def Search(self, num_iterations):
        p = self.mcts_parameters
        # a fresh tree per decision as in CrossingTestEnv::Search; a new Philox range each time
        p.seed = self.mcts_parameters.seed
        self.mcts = Mcts(self.engine, p)
        self.mcts._next_rollout = self._searches << 32
        self._searches += 1
        self.mcts.Search(self.state.batch, num_iterations)
        self.SetJt(self.mcts.ReturnBestAction(merge=False))
        return self.jt

    def SetJt(self, jt):
        self.jt = list(jt)
        self.action_history.append(list(jt))

    def GetJt(self):
        return self.jt

    def GetEgoQval(self):
        """action -> expected reward vector of the ego's root node."""
        st = self.mcts.GetRootStats()[0]
        return {a: st[a, 1:] / st[a, 0] for a in range(st.shape[0]) if st[a, 0] > 0}


@dataclass
class Result:
    """TestRunner::Result (gridworld/test_runner.h:29-43)."""
    pos: int = 0
    value: float = 0.0
    collision: bool = False
    violation: bool = False

    @staticmethod
    def WriteHeader():
        return "Traveled distance\tBase reward\tCollision\tRule violation"

    def __str__(self):
        b = lambda x: "true" if x else "false"
        return "%d\t%g\t%s\t%s" % (self.pos, self.value, b(self.collision), b(self.violation))


class TestRunner:
    __test__ = False  # not a pytest class

    def __init__(self, test_env, q_val_fname=None):
        self.latest_test_env_ = test_env
        self.q_val_fname_ = q_val_fname

    def GetLatestTestEnv(self):
        return self.latest_test_env_

    def GetStateVector(self):
        return self.latest_test_env_.state.GetAgentStates()[:, 0].copy()

    def RunTest(self, num_iter, max_steps=40):
        env = self.latest_test_env_
        steps = 0
        env.state_history.append(self.GetStateVector())
        qw = open(self.q_val_fname_, "w") if self.q_val_fname_ else None
        while not env.state.IsTerminal() and steps < max_steps:
            env.Search(num_iter)
            if qw:  # mvmcts::evaluation::QValWriter: one line per decision, q-values of the ego then its action
                q = env.GetEgoQval()
                qw.write("\t".join("%g" % x for a in sorted(q) for x in q[a]) + "\t%d\n" % env.GetJt()[0])
            step_reward = []
            env.state = env.state.Execute(env.GetJt(), step_reward)
            env.state.ResetDepth()
            env.rewards += np.array(step_reward)
            env.state_history.append(self.GetStateVector())
            steps += 1
        if qw:
            qw.close()
        env.rewards += np.array(env.state.GetTerminalReward())
        ego = env.rewards[0]
        return Result(pos=int(env.state_history[-1][0]), value=float(ego[-1]),
                      collision=bool(ego[0] < 0), violation=bool(ego[1] < 0))
