"""Planner shell (planner.py) on the GPU: the behavioural pins of the reference's
SingleAgentSuite (test/single_agent_test.cpp:229-344) on a world shaped like bark's
make_test_world: accelerate on a free road (best action == 1), brake behind a slower agent
(best action == 4), reach the goal in closed loop without a collision, change_lane."""
import numpy as np
import pytest

import planner_rules_mcts_b200 as brm
from planner_rules_mcts_b200 import planner, scenarios
from planner_rules_mcts_b200.ltl import RuleMonitor

pytestmark = pytest.mark.gpu

# test/single_agent_test.cpp:191-200
PRIMITIVES = [(0.0, 0.0), (5.0, 0.0), (0.0, -1.0), (0.0, 1.0), (-3.0, 0.0)]


def suite_world(n_others, rel_distance, velocity_difference, goal):
    p = brm.RulesMctsStateParameters(GOAL_REWARD=100.0)   # :186-188; everything else default (V = 1)
    return scenarios.straight_road_test_world(
        n_others=n_others, rel_distance=rel_distance, ego_velocity=5.0, velocity_difference=velocity_difference,
        primitives=PRIMITIVES, goal=goal, params=p, rules=[RuleMonitor.MakeRule("G !collision_ego", -1000.0, 0)])


# the reference's suite is typed on BehaviorRulesMctsUct and BehaviorRulesMctsGreedy (:225-227)
STATISTICS = ["uct", "greedy"]


@pytest.mark.parametrize("statistic", STATISTICS)
def test_no_agent_in_front_accelerate(built, statistic):
    scn = suite_world(0, 7.0, 0.0, goal=(7.5, 12.5, -4.25, 0.75))
    pl = planner.BehaviorRulesMcts(scn, planner.PlannerParameters(MaxNumIterations=1000, Statistic=statistic))  # :187
    a = pl.Plan(pl.engine.initial_batch([0]))
    q = pl.GetExpectedRewards()
    assert a == 1 and pl.GetLastAction() == 1, q   # :248
    assert pl.last_stats[0, :, 0].sum() == 1000    # one root visit per iteration
    assert pl.last_stats[0, 1, 0] == pl.last_stats[0, :, 0].max()  # UCT concentrates on the best action


@pytest.mark.parametrize("statistic", STATISTICS)
def test_agent_in_front_must_brake(built, statistic):
    scn = suite_world(1, 2.0, 2.0, goal=(7.5, 12.5, -4.25, 0.75))
    pl = planner.BehaviorRulesMcts(scn, planner.PlannerParameters(MaxNumIterations=1000, Statistic=statistic))
    a = pl.Plan(pl.engine.initial_batch([0]))
    assert a == 4, pl.GetExpectedRewards()         # :268
    # the flat root-parallel search must at least rank braking above accelerating here
    fl = planner.BehaviorRulesMcts(scn, planner.PlannerParameters(MaxNumIterations=20000, Search="flat"),
                                   engine=pl.engine)
    fl.Plan(pl.engine.initial_batch([0]))
    qf = fl.GetExpectedRewards()
    assert qf[4][0] > qf[1][0]


@pytest.mark.parametrize("statistic", STATISTICS)
def test_agent_in_front_reach_goal_closed_loop(built, tmp_path, statistic):
    """:271-305: behind a slower agent (bumper gap 2 m, 2 m/s slower) the planner must bring the ego
    into the goal box ahead (the suite's 5 x 5 polygon moved to (10, -1.75)) without ever touching
    the other vehicle -- with all five motion primitives of the suite, steering included."""
    scn = suite_world(1, 2.0, 2.0, goal=(7.5, 12.5, -4.25, 0.75))
    pl = planner.BehaviorRulesMcts(scn, planner.PlannerParameters(MaxNumIterations=1000, Statistic=statistic))   # :187
    hist, final = planner.run_closed_loop(pl, max_steps=100)                               # :292
    assert final.terminal[0] == 1
    x, y, flags = final.agents[0, 0, 0], final.agents[0, 0, 1], final.flags[0, 0]
    assert flags & 2, "the ego must not have collided"             # EXPECT_FALSE(collision_ego), :295
    assert 7.5 <= x - 1.0 and x + 3.0 <= 12.5 and -4.25 < y < 0.75, "the ego must end inside the goal"  # :296-299
    gap = np.hypot(hist[:, 1, 0] - hist[:, 0, 0], hist[:, 1, 1] - hist[:, 0, 1])
    assert (gap > 2.0).all()  # never closer than two half widths
    path = tmp_path / "states.tsv"
    planner.write_state_history_tsv(str(path), hist)
    lines = path.read_text().splitlines()
    assert lines[0] == "Timestep\tAgent 0\tAgent 1" and len(lines) == hist.shape[0] + 1


@pytest.mark.parametrize("statistic", STATISTICS)
def test_change_lane(built, statistic):
    """:307-344: no other agent, goal box moved to (10, -2) with heading limits +-0.2 rad; the goal
    must be reached within 20 steps."""
    scn = suite_world(0, 2.0, 2.0, goal=(7.5, 12.5, -4.5, 0.5, -0.2, 0.2))
    pl = planner.BehaviorRulesMcts(scn, planner.PlannerParameters(MaxNumIterations=1000, Statistic=statistic))
    hist, final = planner.run_closed_loop(pl, max_steps=20)                                # :330
    x, y, th = final.agents[0, 0, :3]
    assert final.terminal[0] == 1 and final.flags[0, 0] & 2
    assert 7.5 <= x - 1.0 and x + 3.0 <= 12.5 and -4.5 < y < 0.5 and abs(th) < 0.2        # AtGoal, :337-341


def test_terminal_root_is_refused(built):
    scn = suite_world(0, 7.0, 0.0, goal=(-5.0, 5.0, -4.25, 0.75))  # the ego starts inside its goal
    pl = planner.BehaviorRulesMcts(scn)
    with pytest.raises(RuntimeError):
        pl.Plan(pl.engine.initial_batch([0]))


@pytest.mark.parametrize("batch", [8, 32])
def test_leaf_parallel_rounds_keep_the_pins(built, batch):
    """mcts.py with batch_size > 1: several descents per round (virtual loss), one batched Execute
    and one rollout launch for all of them -- same decisions, a fraction of the round trips."""
    for n_others, rel, dv, want in ((0, 7.0, 0.0, 1), (1, 2.0, 2.0, 4)):
        scn = suite_world(n_others, rel, dv, goal=(7.5, 12.5, -4.25, 0.75))
        pl = planner.BehaviorRulesMcts(scn, planner.PlannerParameters(MaxNumIterations=1000, BatchSize=batch))
        assert pl.Plan(pl.engine.initial_batch([0])) == want, pl.GetExpectedRewards()
        assert pl.last_stats[0, :, 0].sum() == 1000            # every descent is counted exactly once
