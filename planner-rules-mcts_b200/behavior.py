"""The stateful planner plugin around the device engine.

`BehaviorRulesMcts` here keeps the bookkeeping of the reference's plugin between planning
steps (/root/reference/src/behavior_rules_mcts.hpp):
  * `Plan(delta_time, observed_world)` (:158-255): the nearest agents interact (they become MCTS
    agents), everybody else is predicted with constant acceleration; rule states of new agents
    are created, those of agents that left are dropped, the automata are carried from the last
    planning step to the current world (`EvaluateRules`), the search runs and the ego's best
    motion primitive is returned as a trajectory;
  * `MakeRuleStates` (:295-350), `GetNewAgents` (:369-385), `GetAgentIdMap` (:392-414),
    `AddKnownAgents` (:415-419), `RemoveRuleStates` (:469-502) with the reference's semantics
    (agent-specific rules are instantiated once per other known agent; the ego is index 0).
BARK's `ObservedWorld` is replaced by a small value class (valid agents by id, lane centre lines,
road polygon).  Every planning step configures the device engine for the agents that are present
(the rule-state layout changes with them) -- a few kilobytes of tables; the search itself is the
host tree of mcts.py over `Execute` / rollouts on the GPU.
"""
import copy
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

from . import _abi
from .continuous import AgentModel, Lane, RulesMctsEngine, RulesMctsStateParameters, Scenario
from .mcts import Mcts, MctsParameters
from .planner import PlannerParameters


@dataclass
class ObservedWorld:
    """What Plan() reads from bark's ObservedWorld: the valid agents (ego included) and the map."""
    ego_id: int
    agents: Dict[int, np.ndarray]          # id -> (x, y, theta, v)
    lanes: List[Lane]
    road: np.ndarray
    time: float = 0.0

    def GetEgoAgentId(self):
        return self.ego_id

    def GetValidAgents(self):
        return dict(sorted(self.agents.items()))

    def GetValidOtherAgents(self):
        return {i: s for i, s in sorted(self.agents.items()) if i != self.ego_id}

    def CurrentEgoPosition(self):
        return np.asarray(self.agents[self.ego_id][:2], np.float64)

    def GetNearestAgents(self, position, num_agents):
        """The `num_agents` agents closest to `position` (the ego is one of them when the
        position is its own, as with bark's r-tree query)."""
        d = sorted((float(np.hypot(*(np.asarray(s[:2], np.float64) - position))), i) for i, s in self.agents.items())
        return {i: self.agents[i] for _, i in d[:num_agents]}


@dataclass
class PredictionSettings:
    """bark PredictionSettings: motion primitives (a, delta) of the ego and of the interacting
    agents; non-interacting agents keep their speed (BehaviorConstantAcceleration, a = 0)."""
    ego_primitives: List[tuple]
    others_primitives: Optional[List[tuple]] = None
    specific_prediction_agents_: set = field(default_factory=set)


@dataclass
class RuleState:
    """ltl::RuleState as far as the planner looks at it."""
    rule: int            # index into the agent's rule list (common rules first)
    agent_ids: tuple     # the other agent an agent-specific rule is bound to, or ()
    q: int               # automaton state
    violations: int = 0


class BehaviorRulesMcts:
    STATISTIC = "uct"

    def __init__(self, params: PlannerParameters, state_params: RulesMctsStateParameters,
                 prediction_settings: PredictionSettings, common_rules=(), agent_rules=None,
                 multi_agent=True, shape: AgentModel = None, ego_goal=None, device=0):
        self.params = params
        self.state_params_ = state_params
        self.prediction_settings_ = prediction_settings
        self.common_rules_ = list(common_rules)
        self.agent_rules_ = {k: list(v) for k, v in (agent_rules or {}).items()}
        self.multi_agent_ = multi_agent
        self.shape = shape or AgentModel()
        self.ego_goal = ego_goal
        self.known_agents_ = set()
        self.multi_agent_rule_state_: Dict[int, List[RuleState]] = {}
        self.last_action_ = None
        self.last_trajectory_ = None
        self.last_stats = None
        self._device = device
        self._abi_engine = None   # created by the first Plan(): the bookkeeping alone needs no GPU
        self._decision = 0
        self.mcts_ = None

    def __repr__(self):
        return "bark.behavior." + type(self).__name__

    # pickling as in python/python_planner_rules_mcts.cpp:44-70: parameters, rules and prediction
    # settings of a model -- not the search state (root, known agents, rule states)
    def __getstate__(self):
        return dict(params=self.params, state_params=self.state_params_,
                    prediction_settings=self.prediction_settings_, common_rules=self.common_rules_,
                    agent_rules=self.agent_rules_, multi_agent=self.multi_agent_, shape=self.shape,
                    ego_goal=self.ego_goal, device=self._device)

    def __setstate__(self, st):
        self.__init__(st["params"], st["state_params"], st["prediction_settings"], st["common_rules"],
                      st["agent_rules"], st["multi_agent"], st["shape"], st["ego_goal"], st["device"])

    def AddLabels(self, label_names):
        """The label functions are built into the kernels; this checks that the names exist."""
        from .continuous import CW_LABELS
        unknown = [n for n in label_names if n.split("#")[0] not in CW_LABELS]
        if unknown:
            raise KeyError("unknown label function(s) %s (known: %s)" % (unknown, sorted(CW_LABELS)))

    def GetTree(self, value_idx=0):
        """(value, ego polyline) per leaf of the last search tree (:431-460)."""
        return self.mcts_.GetTree(value_idx) if self.mcts_ is not None else []

    # -- rules ------------------------------------------------------------------------------
    def AddCommonRules(self, rules):
        self.common_rules_.extend(rules)

    def AddAgentRules(self, agent_ids, rules):
        for i in agent_ids:
            self.agent_rules_.setdefault(i, []).extend(rules)

    def GetCommonRules(self):
        return self.common_rules_

    def GetAgentRules(self):
        return self.agent_rules_

    def _rules_of(self, agent_id):
        return self.common_rules_ + self.agent_rules_.get(agent_id, [])

    def GetNewAgents(self, agent_map):
        return sorted(set(int(i) for i in agent_map) - self.known_agents_)

    def AddKnownAgents(self, agent_ids):
        self.known_agents_.update(agent_ids)

    def MakeRuleStates(self, new_agent_ids):
        current = sorted(set(new_agent_ids) | self.known_agents_)
        for agent_id in new_agent_ids:
            states = self.multi_agent_rule_state_.setdefault(agent_id, [])
            others = [i for i in current if i != agent_id]
            for k, rule in enumerate(self._rules_of(agent_id)):
                states.extend(self._make_rule_state(k, rule, others))
        # missing permutations of the agents that were already known
        for agent_id in sorted(self.known_agents_):
            states = self.multi_agent_rule_state_[agent_id]
            for k, rule in enumerate(self._rules_of(agent_id)):
                if rule.IsAgentSpecific():
                    states.extend(self._make_rule_state(k, rule, list(new_agent_ids)))

    @staticmethod
    def _make_rule_state(k, rule, others):
        if rule.IsAgentSpecific():
            return [RuleState(k, (o,), rule.table.init) for o in others]
        return [RuleState(k, (), rule.table.init)]

    def RemoveRuleStates(self, current_agents):
        current = set(current_agents)
        for agent_id in list(self.multi_agent_rule_state_):
            if agent_id not in current:       # agent is out of the map
                self.known_agents_.discard(agent_id)
                del self.multi_agent_rule_state_[agent_id]
                continue
            self.multi_agent_rule_state_[agent_id] = [
                rs for rs in self.multi_agent_rule_state_[agent_id] if set(rs.agent_ids) <= current]

    def GetAgentIdMap(self, observed_world):
        ids = [observed_world.GetEgoAgentId()]
        if self.multi_agent_:
            ids += [i for i in observed_world.GetValidOtherAgents()
                    if i not in self.prediction_settings_.specific_prediction_agents_]
        return ids

    # -- one planning step --------------------------------------------------------------------
    def _configure(self, observed_world, agent_ids):
        """Device engine for the agents of this world: MCTS agents first (ego = 0), then the
        agents that are only predicted."""
        order = agent_ids + [i for i in observed_world.GetValidAgents() if i not in agent_ids]
        ps = self.prediction_settings_
        models = []
        for k, i in enumerate(order):
            if k == 0:
                prim = ps.ego_primitives
            elif k < len(agent_ids):
                prim = ps.others_primitives or ps.ego_primitives
            else:
                prim = [(0.0, 0.0)]
            m = copy.copy(self.shape)
            m.primitives = list(prim)
            m.goal = self.ego_goal if k == 0 else None
            models.append(m)
        state = np.array([observed_world.agents[i] for i in order], np.float64)
        # the uploaded rule list: the common rules, then the rules given to single agents; every
        # MCTS agent gets the list of rule indices that apply to it (its own rule list, in order)
        rules, per_agent = list(self.common_rules_), []
        for i in agent_ids:
            idx = list(range(len(self.common_rules_)))
            for r in self.agent_rules_.get(i, []):
                k = next((k for k, u in enumerate(rules) if u is r), None)
                if k is None:
                    rules.append(r)
                    k = len(rules) - 1
                idx.append(k)
            per_agent.append(idx)
        only_common = len(rules) == len(self.common_rules_)
        scn = Scenario(self.state_params_, models, len(agent_ids), observed_world.lanes,
                       np.asarray(observed_world.road), rules, state, None if only_common else per_agent)
        return RulesMctsEngine([scn], engine=self._abi_engine), order, rules, per_agent

    @staticmethod
    def _slots(order, a, rules):
        """(uploaded rule index, bound world agent id or None) of every rule slot of MCTS agent a, in
        the engine's order: rules in order, agent-specific ones once per other agent ascending."""
        out = []
        for k, rule in enumerate(rules):
            if rule.IsAgentSpecific():
                out += [(k, o) for j, o in enumerate(order) if j != a]
            else:
                out.append((k, None))
        return out

    def Plan(self, delta_time, observed_world: ObservedWorld):
        nearby = observed_world.GetNearestAgents(observed_world.CurrentEgoPosition(), 3)
        ps = self.prediction_settings_
        ps.specific_prediction_agents_ = {i for i in observed_world.GetValidOtherAgents() if i not in nearby}
        new_agents = self.GetNewAgents(observed_world.GetValidAgents())
        self.MakeRuleStates(new_agents)
        self.AddKnownAgents(new_agents)
        agent_ids = self.GetAgentIdMap(observed_world)
        self.RemoveRuleStates(list(observed_world.GetValidAgents()))
        if self._abi_engine is None:
            self._abi_engine = _abi.Engine(self._device)
        engine, order, rules, per_agent = self._configure(observed_world, agent_ids)
        M, R = engine.M, engine.R
        q0 = np.zeros((M, R), np.uint8)
        index = []
        for a in range(M):
            # the agent's own rule numbering (common rules, then its own) -> uploaded rule index
            states = {(per_agent[a][rs.rule], rs.agent_ids): rs for rs in self.multi_agent_rule_state_[order[a]]}
            slots = self._slots(order, a, rules)
            row = [states.get((k, () if o is None else (o,))) for k, o in slots]   # None: rule of another agent
            assert len(row) == R
            q0[a] = [rules[k].disabled_state if rs is None else rs.q for rs, (k, _) in zip(row, slots)]
            index.append(row)
        root = engine.make_batch(engine.scenarios[0].initial_state[None], q=q0)
        if root.terminal[0]:
            raise RuntimeError("Current state is terminal! Aborting!")     # :208-209
        # transit the automata from the last planning step to the current world (:211-214)
        root, _ = engine.evaluate_rules(root)
        for a in range(M):
            for s, rs in enumerate(index[a]):
                if rs is not None:
                    rs.q = int(root.q[a, s, 0])
                    rs.violations += int(root.viol[a, s, 0])
        root.viol[:] = 0

        sp = self.state_params_
        mp = MctsParameters(engine.V, discount_factor=sp.DISCOUNT_FACTOR,
                            exploration_constant=self.params.UCTExplorationConstant,
                            threshold=self.params.Threshold,
                            max_rollout_steps=self.params.MaxNumIterationsRandomHeuristic,
                            rollouts_per_expansion=self.params.RolloutsPerExpansion,
                            seed=self.params.RandomSeed + self._decision, statistic=self.STATISTIC,
                            batch_size=self.params.BatchSize)
        self._decision += 1
        mcts = Mcts(engine, mp).Search(root, self.params.MaxNumIterations)
        self.last_stats = mcts.GetRootStats()
        best = mcts.ReturnBestAction()[0]
        self.last_action_ = best
        self.mcts_ = mcts
        self.agent_ids_ = agent_ids

        # the ego's motion primitive over delta_time (ego_model->Plan)
        ego = observed_world.agents[observed_world.ego_id]
        tp = copy.copy(sp)
        tp.PREDICTION_TIME_SPAN = float(delta_time)
        m = copy.copy(self.shape)
        m.primitives = list(ps.ego_primitives)
        one = RulesMctsEngine([Scenario(tp, [m], 1, observed_world.lanes, np.asarray(observed_world.road), [],
                                        np.asarray(ego, np.float64)[None])], engine=self._abi_engine)
        b = one.make_batch(one.scenarios[0].initial_state[None])
        b.terminal[:] = 0   # the primitive is executed whatever CheckTerminal says about the horizon
        nxt, _ = one.execute(b, np.array([[best]], np.uint8))
        t0 = observed_world.time
        traj = np.array([[t0, *map(float, ego)], [t0 + delta_time, *map(float, nxt.agents[0, 0])]])
        self.last_trajectory_ = traj
        return traj

    def GetLastAction(self):
        return self.last_action_

    def GetLastTrajectory(self):
        return self.last_trajectory_

    def GetMultiAgentRuleState(self):
        return self.multi_agent_rule_state_

    def GetKnownAgents(self):
        return sorted(self.known_agents_)


class BehaviorRulesMctsUct(BehaviorRulesMcts):
    """BehaviorRulesMcts<ThresUCTStatistic> (python/python_planner_rules_mcts.cpp:30-32)."""
    STATISTIC = "uct"


class BehaviorRulesMctsGreedy(BehaviorRulesMcts):
    """BehaviorRulesMcts<ThresGreedyStatistic> (python/python_planner_rules_mcts.cpp:72-75)."""
    STATISTIC = "greedy"
