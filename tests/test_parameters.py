"""RulesMctsStateParameters: the defaults of the reference's ParamsPtr constructor
(src/rules_mcts_state_parameters.cpp:36-78) in the Python mirror and in the C ABI struct."""
import os
import re

import pytest

import planner_rules_mcts_b200 as brm

# name in the reference's parameter server -> (attribute, default), src/rules_mcts_state_parameters.cpp:36-78
REFERENCE_DEFAULTS = {
    "CollisionWeight": ("COLLISION_WEIGHT", 0.0),
    "OutOfMapWeight": ("OUT_OF_MAP_WEIGHT", -800.0),
    "PotentialWeight": ("POTENTIAL_WEIGHT", 32.0),
    "AccelerationWeight": ("ACCELERATION_WEIGHT", 0.0),
    "RadialAccelerationWeight": ("RADIAL_ACCELERATION_WEIGHT", 0.0),
    "DesiredVelocityWeight": ("DESIRED_VELOCITY_WEIGHT", -5.0),
    "LaneCenterDeviationWeight": ("LANE_CENTER_WEIGHT", 0.0),
    "RewardVectorSize": ("REWARD_VECTOR_SIZE", 1),
    "PredictionTimeSpan": ("PREDICTION_TIME_SPAN", 0.3),
    "DesiredVelocity": ("DESIRED_VELOCITY", 10.0),
    "Horizon": ("HORIZON", 30),
    "DiscountFactor": ("DISCOUNT_FACTOR", 0.9),
    "GoalReward": ("GOAL_REWARD", 0.0),
    "UseRuleRewardForEgoOnly": ("USE_RULE_REWARD_FOR_EGO_ONLY", False),
}


def test_defaults_are_the_reference_defaults():
    p = brm.RulesMctsStateParameters()
    for key, (attr, want) in REFERENCE_DEFAULTS.items():
        assert getattr(p, attr) == want, key
    assert p.REMOVE_AGENTS_OUT_OF_MAP is False     # bark's World default


def test_defaults_reach_the_abi_struct():
    from planner_rules_mcts_b200 import scenarios
    abi = scenarios.highway_scenario(seed=1).to_abi()
    p = scenarios.highway_scenario(seed=1).params
    q = abi.params
    assert q.collision_weight == p.COLLISION_WEIGHT and q.out_of_map_weight == p.OUT_OF_MAP_WEIGHT
    assert q.potential_weight == p.POTENTIAL_WEIGHT and q.desired_velocity_weight == p.DESIRED_VELOCITY_WEIGHT
    assert q.prediction_time_span == p.PREDICTION_TIME_SPAN and q.desired_velocity == p.DESIRED_VELOCITY
    assert q.horizon == p.HORIZON and q.discount_factor == p.DISCOUNT_FACTOR and q.goal_reward == p.GOAL_REWARD
    assert q.reward_vector_size == p.REWARD_VECTOR_SIZE
    assert q.remove_agents_out_of_map == int(p.REMOVE_AGENTS_OUT_OF_MAP)


REF = "/root/reference/src/rules_mcts_state_parameters.cpp"


@pytest.mark.skipif(not os.path.exists(REF), reason="the reference tree is only present in the build container")
def test_table_above_is_what_the_reference_source_says():
    src = open(REF).read()
    # GetReal/GetInt/GetBool("BehaviorRulesMcts::[StateParameters::]<Name>", "<description>", <default>)
    found = {}
    for m in re.finditer(r'Get(Real|Int|Bool)\(\s*"BehaviorRulesMcts::(?:StateParameters::)?(\w+)",\s*'
                         r'(?:"[^"]*"\s*)+,\s*([-\w.]+)\)', src):
        kind, name, val = m.groups()
        found[name] = (val == "true") if kind == "Bool" else float(val)
    assert set(found) == set(REFERENCE_DEFAULTS), sorted(set(found) ^ set(REFERENCE_DEFAULTS))
    for name, (_, want) in REFERENCE_DEFAULTS.items():
        assert found[name] == want, name
