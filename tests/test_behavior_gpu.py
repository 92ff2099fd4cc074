"""The stateful planner plugin (behavior.py) in a closed loop: rule states are
created for new agents, carried from one planning step to the next by EvaluateRules on the GPU,
and dropped with agents that leave (src/behavior_rules_mcts.hpp:158-255,295-350,469-502)."""
import numpy as np
import pytest

import planner_rules_mcts_b200 as brm
from planner_rules_mcts_b200 import behavior, scenarios
from planner_rules_mcts_b200.ltl import RuleMonitor
from planner_rules_mcts_b200.planner import PlannerParameters

pytestmark = pytest.mark.gpu

EGO = 7
PRIMS = [(0.0, 0.0), (1.0, 0.0), (-1.0, 0.0)]


def make_world(agents, t=0.0):
    base = scenarios.straight_road_test_world()
    return behavior.ObservedWorld(EGO, {i: np.asarray(s, np.float64) for i, s in agents.items()},
                                  base.lanes, base.road, t)


def make_planner(iters=150):
    rules = [RuleMonitor.MakeRule("G !collision_ego", -1000.0, 0),
             RuleMonitor.MakeRule("F near_x#0", -1.0, 0)]          # remembers having been near agent #0
    sp = brm.RulesMctsStateParameters(HORIZON=10, NEAR_DISTANCE=12.0)
    return behavior.BehaviorRulesMcts(
        PlannerParameters(MaxNumIterations=iters, RolloutsPerExpansion=8), sp,
        behavior.PredictionSettings(ego_primitives=PRIMS), common_rules=rules, multi_agent=True,
        shape=brm.AgentModel(**scenarios.CAR))


def test_rule_state_bookkeeping_closed_loop(built):
    pl = make_planner()
    F = pl.common_rules_[1].table
    done = [q for q in range(F.n_states) if F.accepting[q]]       # "has been near" states
    # ego 7; car 3 comes up fast from behind on the other lane (not near yet) and overtakes whatever
    # the ego does; 12 far ahead on the other lane, 20 ahead on the ego's lane
    agents = {EGO: (0.0, -1.75, 0.0, 8.0), 3: (-40.0, -5.25, 0.0, 20.0), 12: (60.0, -5.25, 0.0, 8.0),
              20: (50.0, -1.75, 0.0, 9.0)}
    dt = 0.3
    seen_near_3 = None
    for step in range(22):
        if step == 4:
            agents[31] = (agents[EGO][0] + 90.0, -5.25, 0.0, 9.0)  # a new agent appears
        if step == 8:
            del agents[12]                                         # an agent leaves the map
        w = make_world(agents, step * dt)
        traj = pl.Plan(dt, w)
        assert traj.shape == (2, 5) and traj[1, 0] == pytest.approx((step + 1) * dt)
        ids = pl.agent_ids_
        assert ids[0] == EGO and len(ids) == 3                    # ego + the two nearest others
        assert set(ids[1:]) | pl.prediction_settings_.specific_prediction_agents_ == set(agents) - {EGO}
        # one state per common rule and, for the agent-specific rule, per other known agent
        assert pl.GetKnownAgents() == sorted(agents)
        for i in agents:
            rs = pl.GetMultiAgentRuleState()[i]
            assert sorted((r.rule, r.agent_ids) for r in rs) == \
                sorted([(0, ())] + [(1, (o,)) for o in agents if o != i])
        ego_rs = {r.agent_ids: r for r in pl.GetMultiAgentRuleState()[EGO] if r.rule == 1}
        gap = agents[3][0] - agents[EGO][0]
        if seen_near_3 is None and ego_rs[(3,)].q in done:
            seen_near_3 = step
            assert abs(gap) < 12.0 + 1e-3
        if seen_near_3 is not None:
            assert ego_rs[(3,)].q in done                          # carried across planning steps
        elif abs(gap) > 12.0 + 1e-3:
            assert ego_rs[(3,)].q == F.init
        # world step: ego follows its trajectory, the others keep their speed
        agents[EGO] = tuple(traj[1, 1:])
        for i in agents:
            if i != EGO:
                x, y, th, v = agents[i]
                agents[i] = (x + v * dt, y, th, v)
    assert seen_near_3 is not None                                 # agent 3 came near ...
    assert agents[3][0] - agents[EGO][0] > 12.0                    # ... and is far ahead again: the state was kept
    assert agents[20][0] - agents[EGO][0] > 4.0                    # nobody was run into
    assert 12 not in pl.GetMultiAgentRuleState() and 31 in pl.GetMultiAgentRuleState()


def test_terminal_root_is_refused(built):
    pl = make_planner(10)
    w = make_world({EGO: (0.0, -1.75, 0.0, 8.0), 3: (2.0, -1.75, 0.0, 3.0)})   # already overlapping
    with pytest.raises(RuntimeError):
        pl.Plan(0.3, w)


def test_evaluate_rules_matches_execute_transition(built):
    """brm_cw_evaluate_rules on Execute's successor kinematics with the predecessor's automaton
    states gives exactly Execute's automaton states (no collision in the step)."""
    scn = scenarios.merge_scenario()
    eng = brm.RulesMctsEngine([scn])
    b = eng.initial_batch([0]).select(np.zeros(64, np.int64))
    rng = np.random.default_rng(5)
    for _ in range(6):
        acts = rng.integers(0, 3, size=(eng.M, b.n)).astype(np.uint8)
        nxt, _ = eng.execute(b, acts)
        probe = brm.CwBatch(nxt.scenario.copy(), nxt.agents.copy(), nxt.flags.copy(), nxt.horizon.copy(),
                            nxt.terminal.copy(), b.q.copy(), np.zeros_like(b.viol))
        out, rew = eng.evaluate_rules(probe)
        same_flags = (nxt.flags == b.flags).all(axis=0)            # nobody was invalidated in this step
        assert same_flags.any()
        assert (out.q[:, :, same_flags] == nxt.q[:, :, same_flags]).all()
        assert (out.agents == nxt.agents).all() and (out.horizon == nxt.horizon).all()
        assert (rew <= 0).all()
        keep = ~nxt.terminal.astype(bool)
        if not keep.any():
            break
        b = nxt.select(np.flatnonzero(keep))


def test_python_binding_surface(built):
    """python/python_planner_rules_mcts.cpp:37-41,113-122: GetTree, Execute -> (state, rewards),
    observed_world."""
    pl = make_planner(60)
    assert pl.GetTree() == []
    w = make_world({EGO: (0.0, -1.75, 0.0, 8.0), 3: (40.0, -1.75, 0.0, 6.0)})
    pl.Plan(0.3, w)
    tree = pl.GetTree(0)
    assert len(tree) >= 3                                          # at least the root's children
    for value, line in tree:
        assert np.isfinite(value) and len(line) >= 2
        assert line[0] == pytest.approx((0.0, -1.75))              # every polyline starts at the ego
        assert all(b[0] >= a[0] for a, b in zip(line, line[1:]))   # and drives forward
    scn = scenarios.straight_road_test_world()
    s0 = brm.MvmctsState(brm.RulesMctsEngine([scn]))
    s1, rewards = s0.Execute([0])
    assert isinstance(s1, brm.MvmctsState) and len(rewards) == 1 and rewards[0].shape == (1,)
    states, flags = s1.observed_world
    assert states.shape == (2, 4) and (flags == 3).all()
    assert [r.tolist() for r in s1.EvaluateRules()] == [[0.0]]


def test_rules_given_to_single_agents(built):
    """AddAgentRules (:278-293): the ego alone carries the "has been near" rule; the other interacting
    agents are searched with the common rule only, and the ego's state still advances."""
    rules = [RuleMonitor.MakeRule("G !collision_ego", -1000.0, 0)]
    near = RuleMonitor.MakeRule("F near_x#0", -1.0, 0)
    sp = brm.RulesMctsStateParameters(HORIZON=10, NEAR_DISTANCE=12.0)
    pl = behavior.BehaviorRulesMctsUct(PlannerParameters(MaxNumIterations=100, RolloutsPerExpansion=8), sp,
                                       behavior.PredictionSettings(ego_primitives=PRIMS), common_rules=rules,
                                       shape=brm.AgentModel(**scenarios.CAR))
    pl.AddAgentRules([EGO], [near])
    agents = {EGO: (0.0, -1.75, 0.0, 8.0), 3: (8.0, -5.25, 0.0, 8.0), 12: (60.0, -5.25, 0.0, 8.0)}
    pl.Plan(0.3, make_world(agents))
    st = pl.GetMultiAgentRuleState()
    assert sorted((r.rule, r.agent_ids) for r in st[EGO]) == [(0, ()), (1, (3,)), (1, (12,))]
    assert sorted((r.rule, r.agent_ids) for r in st[3]) == [(0, ())]
    ego = {r.agent_ids: r for r in st[EGO] if r.rule == 1}
    done = [q for q in range(near.table.n_states) if near.table.accepting[q]]
    assert ego[(3,)].q in done and ego[(12,)].q == near.table.init       # 3 is near, 12 is not
