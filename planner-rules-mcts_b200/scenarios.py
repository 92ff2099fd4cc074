"""Synthetic scenarios of the shapes named in BASELINE.json (configs S, M, W)
and small worlds shaped like the reference's own tests (test/single_agent_test.cpp,
test/multi_agent_test.cpp).  Everything is generated from a seed; there is no map data."""
import math

import numpy as np

from .continuous import BIG, AgentModel, Lane, RulesMctsStateParameters, Scenario
from .gridworld import ZIP_FORMULA
from .ltl import RuleMonitor

LANE_WIDTH = 3.5
CAR = dict(front=3.0, rear=1.0, half_width=0.9)  # 4 m vehicle (test/single_agent_test.cpp:94-95)


def _line(p0, p1, n):
    t = np.linspace(0.0, 1.0, n)[:, None]
    return (1 - t) * np.asarray(p0, np.float64) + t * np.asarray(p1, np.float64)


def straight_road_test_world(n_others=1, rel_distance=7.0, ego_velocity=5.0, velocity_difference=0.0,
                             multi_agent=False, primitives=None, goal=(20.0, 27.0, -6.25, 0.75),
                             horizon=20, rules=None, params=None):
    """World shaped like bark's make_test_observed_world as used by
    test/single_agent_test.cpp:121-123 and test/multi_agent_test.cpp:95-103: a long straight
    corridor along x with lane centre y = -1.75, ego at x = 0, others in front."""
    p = params or RulesMctsStateParameters(POTENTIAL_WEIGHT=0.0, DESIRED_VELOCITY_WEIGHT=0.0, HORIZON=horizon)
    if primitives is None:
        dt = p.PREDICTION_TIME_SPAN
        primitives = [(0.0, 0.0), (50.0, 1.0), ((rel_distance + 4.0) * 2.0 / (dt * dt), 0.0)]
    if rules is None:
        rules = [RuleMonitor.MakeRule("G !collision_ego", -1000.0, 0)]
    A = 1 + n_others
    M = A if multi_agent else 1
    agents = []
    for i in range(A):
        prim = primitives if i < M else [(0.0, 0.0)]  # BehaviorConstantAcceleration, a = 0
        agents.append(AgentModel(primitives=list(prim), goal=goal if i == 0 else None, **CAR))
    lanes = [Lane(_line((-50.0, -1.75), (400.0, -1.75), 16), LANE_WIDTH / 2, rightmost=False),
             Lane(_line((-50.0, -5.25), (400.0, -5.25), 16), LANE_WIDTH / 2, rightmost=True)]
    road = np.array([(-50.0, 0.0), (-50.0, -7.0), (400.0, -7.0), (400.0, 0.0)])
    state = np.zeros((A, 4), np.float64)
    state[0] = (0.0, -1.75, 0.0, ego_velocity)
    for i in range(1, A):
        state[i] = (i * (rel_distance + 4.0), -1.75, 0.0, ego_velocity - velocity_difference)
    return Scenario(p, agents, M, lanes, road, rules, state)


def highway_scenario(seed=1000, n_others=3, pts_per_lane=64, steer=0.1):
    """Config S: single-agent highway lane change, safe-distance rule, horizon 20.  Ego is the
    only MCTS agent (BehaviorRulesMcts::MultiAgent = false); the others keep their speed."""
    rng = np.random.default_rng(seed)
    p = RulesMctsStateParameters(REWARD_VECTOR_SIZE=2, HORIZON=20)
    prim = [(0.0, 0.0), (5.0, 0.0), (0.0, -steer), (0.0, steer), (-3.0, 0.0)]  # test/single_agent_test.cpp:191-200
    agents = [AgentModel(primitives=prim, goal=(250.0, 300.0, -7.0, 0.0), **CAR)]
    agents += [AgentModel(primitives=[(0.0, 0.0)], **CAR) for _ in range(n_others)]
    lanes = [Lane(_line((-40.0, -1.75), (340.0, -1.75), pts_per_lane), LANE_WIDTH / 2),
             Lane(_line((-40.0, -5.25), (340.0, -5.25), pts_per_lane), LANE_WIDTH / 2, rightmost=True)]
    road = np.array([(-40.0, 0.0), (-40.0, -7.0), (340.0, -7.0), (340.0, 0.0)])
    rules = [RuleMonitor.MakeRule("G safe_distance_front", -1.0, 0),
             RuleMonitor.MakeRule("G !collision_ego", -1000.0, 0)]
    state = np.zeros((1 + n_others, 4), np.float64)
    state[0] = (0.0, -1.75, 0.0, 5.0)
    x = 0.0
    for i in range(1, 1 + n_others):
        x += rng.uniform(7.0, 40.0) + 4.0
        lane_y = -1.75 if rng.random() < 0.6 else -5.25
        state[i] = (x, lane_y, 0.0, rng.uniform(3.0, 10.0))
    return Scenario(p, agents, 1, lanes, road, rules, state)


def merge_scenario(seed=1000, n_agents=4, pts_per_lane=32, primitives=None):
    """Config M: two lanes merging into one, all agents controlled (multi_agent_test shape),
    zipper + safe-distance rules, V = 3 (0 safe distance, 1 zipper, 2 comfort), horizon 20."""
    rng = np.random.default_rng(seed)
    p = RulesMctsStateParameters(REWARD_VECTOR_SIZE=3, HORIZON=20)
    prim = primitives or [(0.0, 0.0), (-3.0, 0.0), (4.0, 0.0)]
    # The ramp is a straight line that meets the main lane at the merge point under a shallow
    # angle (3.5 m lateral per 100 m), so agents without steering primitives stay on their
    # lane; past the merge point ramp agents drift slowly across the merged lane, the road has
    # a shoulder (y up to +1.75) to keep them inside for the length of a horizon.
    merge_x, ramp_len = 100.0, 260.0
    ang = math.atan2(LANE_WIDTH, 100.0)
    ramp0 = (merge_x - ramp_len * math.cos(ang), -1.75 - ramp_len * math.sin(ang))
    lanes = [
        Lane(_line((merge_x, -1.75), (330.0, -1.75), pts_per_lane), LANE_WIDTH / 2, s_point=-BIG),       # merged
        Lane(_line((-160.0, -1.75), (merge_x, -1.75), pts_per_lane), LANE_WIDTH / 2, successor=0),       # main
        Lane(_line(ramp0, (merge_x, -1.75), pts_per_lane), LANE_WIDTH / 2, successor=0, rightmost=True),  # ramp
    ]
    ramp_y = lambda x: -1.75 - (merge_x - x) * math.tan(ang)
    x_join = merge_x + 0.25 / math.tan(ang)  # where the ramp's lower road edge reaches y = -3.5
    road = np.array([(-160.0, 1.75), (-160.0, ramp_y(-160.0) - 2.0), (x_join, -3.5), (330.0, -3.5),
                     (330.0, 1.75)])
    agents = [AgentModel(primitives=list(prim), goal=(280.0, 330.0, -3.5, 1.75), **CAR) for _ in range(n_agents)]
    rules = [RuleMonitor.MakeRule("G safe_distance_front", -1.0, 0),
             RuleMonitor.MakeRule(ZIP_FORMULA, -1.0, 1)]
    state = np.zeros((n_agents, 4), np.float64)
    back = [40.0, 62.0]  # distance before the merge point already occupied on main / ramp
    for i in range(n_agents):
        on_ramp = i % 2 == 1
        d = back[on_ramp] + rng.uniform(5.0, 15.0) + 4.0
        back[on_ramp] = d
        v = rng.uniform(9.0, 14.0)
        if on_ramp:
            state[i] = (merge_x - d * math.cos(ang), -1.75 - d * math.sin(ang), ang, v)
        else:
            state[i] = (merge_x - d, -1.75, 0.0, v)
    return Scenario(p, agents, n_agents, lanes, road, rules, state)


def intersection_scenario(seed=1000, scenario_id=0, n_agents=10, pts_per_lane=32):
    """Config W: one synthetic 4-way intersection (4 approach + 4 exit lanes), 10 controlled
    agents, 5 rules (3 ego-only + 2 with a #0 placeholder -> 3 + 2*9 = 21 rule states per
    agent), V = 4, horizon 20.  Geometry and initial states depend on (seed, scenario_id)."""
    rng = np.random.default_rng([seed, scenario_id])
    p = RulesMctsStateParameters(REWARD_VECTOR_SIZE=4, HORIZON=20)
    hw = LANE_WIDTH / 2
    arm = rng.uniform(50.0, 70.0)
    w = LANE_WIDTH  # arm half-width: one approach + one exit lane
    lanes = []
    dirs = [0.0, math.pi / 2, math.pi, 3 * math.pi / 2]  # travel direction of approach k
    for k, th in enumerate(dirs):  # approach lanes 0..3
        c, s = math.cos(th), math.sin(th)
        rot = lambda x, y: (x * c - y * s, x * s + y * c)
        lanes.append(Lane(_line(rot(-arm, -hw), rot(-w, -hw), pts_per_lane), hw, successor=4 + k, s_point=BIG))
    for k, th in enumerate(dirs):  # exit lanes 4..7 = straight across (junction interior + far arm)
        c, s = math.cos(th), math.sin(th)
        rot = lambda x, y: (x * c - y * s, x * s + y * c)
        lanes.append(Lane(_line(rot(-w, -hw), rot(arm, -hw), pts_per_lane), hw, s_point=-BIG))
    road = np.array([(-arm, w), (-arm, -w), (-w, -w), (-w, -arm), (w, -arm), (w, -w), (arm, -w), (arm, w),
                     (w, w), (w, arm), (-w, arm), (-w, w)])
    prim = [(0.0, 0.0), (-3.0, 0.0), (3.0, 0.0)]
    agents = [AgentModel(primitives=list(prim), **CAR) for _ in range(n_agents)]
    rules = [RuleMonitor.MakeRule("G !collision_ego", -1000.0, 0),
             RuleMonitor.MakeRule("G !collision_corridor", -100.0, 0),
             RuleMonitor.MakeRule("G safe_distance_front", -1.0, 1),
             RuleMonitor.MakeRule(ZIP_FORMULA, -1.0, 2),
             RuleMonitor.MakeRule("G !(near_x#0 & in_direct_front_x#0)", -1.0, 2)]
    state = np.zeros((n_agents, 4), np.float64)
    back = [w + 3.0] * 4
    for i in range(n_agents):
        k = i % 4
        d = back[k] + rng.uniform(3.0, 9.0)
        back[k] = d + 4.0
        th = dirs[k]
        c, s = math.cos(th), math.sin(th)
        # a lateral offset from the centre line per agent: two agents exactly on one line would
        # project to the SAME arc length on every lane that line crosses, and "who is in front"
        # (gap > 0) would be decided by rounding noise -- in the reference's double arithmetic too
        x, y = -d, -hw + rng.uniform(-0.35, 0.35)
        state[i] = (x * c - y * s, x * s + y * c, th, rng.uniform(2.0, 7.0))
    return Scenario(p, agents, n_agents, lanes, road, rules, state)
