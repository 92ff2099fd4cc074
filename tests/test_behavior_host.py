"""Host bookkeeping of the planner plugin (behavior.py) against the reference's semantics
(src/behavior_rules_mcts.hpp:295-350 MakeRuleStates, :369-385 GetNewAgents, :392-414
GetAgentIdMap, :469-502 RemoveRuleStates).  No GPU involved."""
import numpy as np

import planner_rules_mcts_b200 as brm
from planner_rules_mcts_b200 import behavior
from planner_rules_mcts_b200.ltl import RuleMonitor
from planner_rules_mcts_b200.planner import PlannerParameters


def planner(multi_agent=True):
    rules = [RuleMonitor.MakeRule("G !collision_ego", -1000.0, 0),
             RuleMonitor.MakeRule("G (in_direct_front_x#0 -> !near_x#0)", -1.0, 1)]
    return behavior.BehaviorRulesMcts(PlannerParameters(), brm.RulesMctsStateParameters(),
                                      behavior.PredictionSettings(ego_primitives=[(0.0, 0.0)]),
                                      common_rules=rules, multi_agent=multi_agent)


def keys(pl, agent):
    return sorted((r.rule, r.agent_ids) for r in pl.GetMultiAgentRuleState()[agent])


def test_make_rule_states_for_new_and_known_agents():
    pl = planner()
    new = pl.GetNewAgents({5: None, 2: None, 9: None})
    assert new == [2, 5, 9]                                    # sorted ids, all unknown
    pl.MakeRuleStates(new)
    pl.AddKnownAgents(new)
    # one state of the common rule, one state of the agent-specific rule per other agent
    assert keys(pl, 5) == [(0, ()), (1, (2,)), (1, (9,))]
    assert keys(pl, 2) == [(0, ()), (1, (5,)), (1, (9,))]
    init = pl.common_rules_[1].table.init
    assert all(r.q == (init if r.rule == 1 else pl.common_rules_[0].table.init)
               for r in pl.GetMultiAgentRuleState()[9])
    # a later agent: its own states against everybody, and the missing permutation for the known ones
    pl.GetMultiAgentRuleState()[5][1].q = 3                     # some progress that must survive
    new = pl.GetNewAgents({2: None, 5: None, 9: None, 7: None})
    assert new == [7]
    pl.MakeRuleStates(new)
    pl.AddKnownAgents(new)
    assert keys(pl, 7) == [(0, ()), (1, (2,)), (1, (5,)), (1, (9,))]
    assert keys(pl, 5) == [(0, ()), (1, (2,)), (1, (7,)), (1, (9,))]
    assert pl.GetMultiAgentRuleState()[5][1].q == 3
    assert pl.GetNewAgents({2: None, 5: None}) == []


def test_remove_rule_states_of_agents_that_left():
    pl = planner()
    pl.MakeRuleStates([1, 2, 3])
    pl.AddKnownAgents([1, 2, 3])
    pl.RemoveRuleStates([1, 3])
    assert pl.GetKnownAgents() == [1, 3] and 2 not in pl.GetMultiAgentRuleState()
    assert keys(pl, 1) == [(0, ()), (1, (3,))]                 # states bound to agent 2 are gone
    assert keys(pl, 3) == [(0, ()), (1, (1,))]
    # the agent comes back: it is new again and gets fresh states
    assert pl.GetNewAgents({1: None, 2: None, 3: None}) == [2]


def test_agent_id_map_and_nearest_agents():
    lanes, road = [], np.zeros((4, 2))
    agents = {4: np.array([0.0, 0.0, 0.0, 5.0]), 1: np.array([50.0, 0.0, 0.0, 5.0]),
              8: np.array([-10.0, 3.5, 0.0, 5.0]), 6: np.array([20.0, 0.0, 0.0, 5.0])}
    w = behavior.ObservedWorld(4, agents, lanes, road)
    near = w.GetNearestAgents(w.CurrentEgoPosition(), 3)
    assert set(near) == {4, 8, 6}                              # the ego itself counts, as with bark's r-tree query
    pl = planner()
    pl.prediction_settings_.specific_prediction_agents_ = {i for i in w.GetValidOtherAgents() if i not in near}
    assert pl.prediction_settings_.specific_prediction_agents_ == {1}
    assert pl.GetAgentIdMap(w) == [4, 6, 8]                    # ego first, interacting others by id
    assert planner(multi_agent=False).GetAgentIdMap(w) == [4]


def test_agent_rules_are_kept_per_agent():
    pl = planner()
    extra = RuleMonitor.MakeRule("G safe_distance_front", -1.0, 0)
    pl.AddAgentRules([3], [extra])
    pl.MakeRuleStates([3, 4])
    assert keys(pl, 3) == [(0, ()), (1, (4,)), (2, ())] and keys(pl, 4) == [(0, ()), (1, (3,))]
    assert pl.GetAgentRules()[3] == [extra] and len(pl.GetCommonRules()) == 2


def test_pickle_round_trip_of_an_unused_model():
    """python/python_planner_rules_mcts.cpp:44-70: parameters, rules and prediction settings."""
    import pickle
    pl = behavior.BehaviorRulesMctsGreedy(PlannerParameters(MaxNumIterations=77), brm.RulesMctsStateParameters(HORIZON=9),
                                          behavior.PredictionSettings(ego_primitives=[(0.0, 0.0), (1.0, 0.1)]),
                                          common_rules=[RuleMonitor.MakeRule("G !collision_ego", -1000.0, 0)])
    pl.AddAgentRules([4], [RuleMonitor.MakeRule("G safe_distance_front", -1.0, 0)])
    pl.MakeRuleStates([1, 4])
    q = pickle.loads(pickle.dumps(pl))
    assert type(q) is behavior.BehaviorRulesMctsGreedy and repr(q) == "bark.behavior.BehaviorRulesMctsGreedy"
    assert q.params.MaxNumIterations == 77 and q.state_params_.HORIZON == 9 and q.STATISTIC == "greedy"
    assert q.prediction_settings_.ego_primitives == [(0.0, 0.0), (1.0, 0.1)]
    assert [r.table.trans for r in q.GetCommonRules()] == [r.table.trans for r in pl.GetCommonRules()]
    assert list(q.GetAgentRules()) == [4] and q.GetAgentRules()[4][0].weight == -1.0
    assert q.GetMultiAgentRuleState() == {} and q.GetKnownAgents() == []   # the search state is not pickled
    q.AddLabels(["near_x#0", "collision_ego"])
    import pytest
    with pytest.raises(KeyError):
        q.AddLabels(["no_such_label"])
    assert q.GetTree() == []
