"""The device arithmetic of the continuous-world kernels (csrc/cw_core.cuh, compiled for the host
by tests/cpp/cw_emul.cpp and run work item by work item) against the FP64 oracle -- the part of
the GPU parity suite that can run without a GPU.  Same policy as tests/test_cw_gpu.py: every
discrete output EQUAL on every trace / rollout, floats within 1e-9 relative."""
import numpy as np
import pytest

from oracle import cw as ocw
from planner_rules_mcts_b200 import scenarios
from planner_rules_mcts_b200.continuous import AgentModel, RulesMctsStateParameters, Scenario
from planner_rules_mcts_b200.ltl import RuleMonitor
from tests.cw_common import assert_close, compare_rollouts, compare_trace, oracle_config
from tests.cw_emul import CwEmul

FL = lambda n: np.full(n, ocw.PRESENT | ocw.VALID, np.uint8)


@pytest.fixture(scope="module")
def orc(built):
    return ocw.CwOracle()


@pytest.fixture(scope="module")
def em():
    return CwEmul()


def weighted(scn):
    p = scn.params
    p.ACCELERATION_WEIGHT, p.RADIAL_ACCELERATION_WEIGHT, p.LANE_CENTER_WEIGHT = -0.5, -1.0, -2.0
    p.GOAL_REWARD, p.COLLISION_WEIGHT = 100.0, -50.0
    return scn


SCN = {
    "highway": lambda: weighted(scenarios.highway_scenario(seed=5)),
    "merge": lambda: weighted(scenarios.merge_scenario(seed=5)),
    "intersection": lambda: weighted(scenarios.intersection_scenario(seed=5)),
    "test_world": lambda: weighted(scenarios.straight_road_test_world(
        n_others=2, multi_agent=True, primitives=[(0.0, 0.0), (3.0, 0.05), (-2.0, -0.05), (4.0, 0.0)])),
}


def test_sincos_is_within_two_ulp(em):
    rng = np.random.default_rng(0)
    xs = np.concatenate([rng.uniform(-12, 12, 20000), rng.uniform(-1e-3, 1e-3, 2000), [0.0, np.pi / 2, -np.pi, 3.0]])
    worst = 0.0
    for x in xs:
        s, c = em.sincos(float(x))
        for got, want in ((s, np.sin(x)), (c, np.cos(x))):
            worst = max(worst, abs(got - want) / np.spacing(abs(want)))
    assert worst <= 2.0, worst


@pytest.mark.parametrize("name", sorted(SCN))
def test_lane_projection_is_the_brute_force_search(em, orc, name):
    """cw_frenet (culled, FP64) vs the oracle's search over every segment, via one-step replays."""
    scn = SCN[name]()
    cfg = oracle_config(scn)
    rng = np.random.default_rng(3)
    pts = np.asarray(scn.road)
    lo, hi = pts.min(0) - 5.0, pts.max(0) + 5.0
    A = len(scn.agents)
    n_on = 0
    for _ in range(400):
        x, y = rng.uniform(lo, hi)
        lane, s, lat = em.frenet(scn, float(x), float(y))
        # the oracle projects agents inside Execute: put a motionless agent at (x, y)
        st = scn.initial_state.copy()
        st[A - 1] = (x, y, 0.3, 0.0)
        fl = FL(A)
        fl[A - 1] = ocw.PRESENT                      # INVALID: frozen, but projected
        o = orc.replay(cfg, st, fl, 5, np.zeros((1, scn.n_mcts_agents), np.uint8))
        if o["steps"] == 0:
            continue
        want = o["frenet"][0, A - 1]
        assert lane == int(want[0]), (x, y, lane, want)
        np.testing.assert_allclose([s, lat], want[1:], rtol=1e-12, atol=1e-12)
        n_on += lane >= 0
    assert n_on > 20


@pytest.mark.parametrize("name", sorted(SCN))
def test_lane_projection_at_the_band_edges(em, orc, name):
    """Points within a hair of a lane's half width, of its ends and of its segment joints: where a
    culled search would go wrong first."""
    scn = SCN[name]()
    cfg = oracle_config(scn)
    rng = np.random.default_rng(5)
    A = len(scn.agents)
    checked = 0
    for lane in scn.lanes:
        pts = np.asarray(lane.points, np.float64)
        for _ in range(40):
            k = rng.integers(0, len(pts) - 1)
            t = rng.choice([0.0, 1.0, rng.random(), 1e-9, 1 - 1e-9, -1e-7, 1 + 1e-7])
            d = pts[k + 1] - pts[k]
            nrm = np.array([-d[1], d[0]]) / np.hypot(*d)
            off = rng.choice([1.0, -1.0]) * lane.half_width * rng.choice([1.0, 1 - 1e-9, 1 + 1e-9, 1 - 1e-4, 1 + 1e-4, 0.5])
            x, y = pts[k] + t * d + off * nrm
            got = em.frenet(scn, float(x), float(y))
            st = scn.initial_state.copy()
            st[A - 1] = (x, y, 0.3, 0.0)
            fl = FL(A)
            fl[A - 1] = ocw.PRESENT
            o = orc.replay(cfg, st, fl, 5, np.zeros((1, scn.n_mcts_agents), np.uint8))
            if o["steps"] == 0:
                continue
            want = o["frenet"][0, A - 1]
            assert got[0] == int(want[0]), (x, y, got, want)
            np.testing.assert_allclose(got[1:], want[1:], rtol=1e-12, atol=1e-12)
            checked += 1
    assert checked > 20


@pytest.mark.parametrize("name", sorted(SCN))
def test_drivable_area_grid_never_disagrees_with_the_exact_test(em, name):
    """cw_classify_shape (fp32 grid over the road polygon) either decides like cw_out_of_map or
    leaves the pose to it; poses all over the map and hugging every edge and vertex."""
    scn = SCN[name]()
    road = np.asarray(scn.road, np.float64)
    rng = np.random.default_rng(9)
    lo, hi = road.min(0) - 6.0, road.max(0) + 6.0
    counts = np.zeros(4, int)
    for _ in range(1500):
        x, y = rng.uniform(lo, hi)
        counts[em.classify(scn, float(x), float(y), float(rng.uniform(-np.pi, np.pi)))] += 1
    for e in range(len(road)):
        a, b = road[e], road[(e + 1) % len(road)]
        for _ in range(60):
            p = a + rng.random() * (b - a) + rng.normal(0, rng.choice([1e-3, 0.05, 1.5]), 2)
            counts[em.classify(scn, float(p[0]), float(p[1]), float(rng.uniform(-np.pi, np.pi)))] += 1
    decided = counts[0] + counts[1]
    assert counts[0] > 0 and counts[1] > 0 and decided > 0.5 * counts.sum(), counts


@pytest.mark.parametrize("name", sorted(SCN))
def test_replay_matches_oracle(em, orc, name):
    scn = SCN[name]()
    cfg = oracle_config(scn)
    rng = np.random.default_rng(11)
    A, M = len(scn.agents), scn.n_mcts_agents
    n_act = len(scn.agents[0].primitives)
    total, margin = 0, np.inf
    for b in range(64):
        acts = rng.integers(0, n_act, size=(22, M)).astype(np.uint8)
        o = orc.replay(cfg, scn.initial_state, FL(A), scn.params.HORIZON, acts)
        tr = em.replay(scn, scn.initial_state, FL(A), scn.params.HORIZON, acts)
        total += compare_trace(tr, 0, o, scn)
        if o["steps"]:
            margin = min(margin, o["margin"][:o["steps"]].min())
    assert total > 3 * 64 and margin > 1e-9


def test_replay_perturbed_states_and_flags(em, orc):
    """Random start states incl. collided / absent agents and a frozen ego."""
    scn = scenarios.merge_scenario(seed=5)
    cfg = oracle_config(scn)
    rng = np.random.default_rng(2)
    A = 4
    n_term0 = 0
    for b in range(96):
        st = scn.initial_state.copy()
        st[:, 0] += rng.uniform(-15, 25, size=A)
        st[:, 3] += rng.uniform(-4, 4, size=A)
        fl = FL(A)
        fl[rng.random(A) < 0.15] = ocw.PRESENT
        fl[1:][rng.random(A - 1) < 0.05] = 0
        acts = rng.integers(0, 3, size=(12, A)).astype(np.uint8)
        t0 = orc.check_terminal(cfg, st, fl, 9)
        assert em.check_terminal(scn, st, fl, 9) == t0
        n_term0 += t0
        o = orc.replay(cfg, st, fl, 9, acts)
        tr = em.replay(scn, st, fl, 9, acts)
        compare_trace(tr, 0, o, scn)
    assert 0 < n_term0 < 96


@pytest.mark.parametrize("name,n", [("merge", 4096), ("highway", 2048), ("intersection", 512)])
def test_rollouts_match_oracle(em, orc, name, n):
    scn = {"highway": scenarios.highway_scenario, "merge": scenarios.merge_scenario,
           "intersection": scenarios.intersection_scenario}[name](seed=21)
    cfg = oracle_config(scn)
    A = len(scn.agents)
    g = scn.params.DISCOUNT_FACTOR
    ref = orc.rollout_detail(cfg, scn.initial_state, FL(A), scn.params.HORIZON, 77, 1000, n, 30, g)
    out = em.rollout_detail(scn, scn.initial_state, FL(A), scn.params.HORIZON, 77, 1000, n, 30, g)
    compare_rollouts(out, ref, name)
    assert ref["margin"].min() > 1e-9
    agg = orc.rollout(cfg, scn.initial_state, FL(A), scn.params.HORIZON, 77, 1000, n, 30, g, n_threads=2)
    assert out["agent_steps"] == agg["agent_steps"]


@pytest.mark.parametrize("n_agents,n_mcts,n", [(1, 1, 333), (2, 1, 100), (7, 7, 65), (16, 16, 50), (16, 3, 31),
                                               (16, 1, 40), (5, 2, 77)])
def test_edge_shapes(em, orc, n_agents, n_mcts, n):
    scn = edge_scenario(n_agents, n_mcts)
    cfg = oracle_config(scn)
    g = scn.params.DISCOUNT_FACTOR
    ref = orc.rollout_detail(cfg, scn.initial_state, FL(n_agents), 12, 5, 0, n, 30, g)
    out = em.rollout_detail(scn, scn.initial_state, FL(n_agents), 12, 5, 0, n, 30, g)
    compare_rollouts(out, ref, (n_agents, n_mcts))
    assert ref["margin"].min() > 1e-9


def edge_scenario(n_agents, n_mcts):
    """One agent, the 16-agent maximum, mixes of controlled and simulated-only agents; many rule slots."""
    rng = np.random.default_rng(n_agents * 17 + n_mcts)
    base = scenarios.straight_road_test_world(n_others=0)
    p = RulesMctsStateParameters(REWARD_VECTOR_SIZE=3, HORIZON=12)
    prim = [(0.0, 0.0), (2.0, 0.0), (-3.0, 0.0), (0.0, 0.03)]
    agents = []
    for i in range(n_agents):
        agents.append(AgentModel(primitives=prim if i < n_mcts else [(0.5, 0.0)], front=3.0, rear=1.0,
                                 half_width=0.9, goal=(150.0, 400.0, -7.0, 0.0) if i == 0 else None))
    rules = [RuleMonitor.MakeRule("G !collision_ego", -100.0, 0),
             RuleMonitor.MakeRule("G safe_distance_front", -1.0, 1),
             RuleMonitor.MakeRule("G !(near_x#0 & in_direct_front_x#0)", -1.0, 1)]
    state = np.zeros((n_agents, 4))
    for i in range(n_agents):
        lane_y = -1.75 if i % 2 == 0 else -5.25
        state[i] = ((i // 2) * 14.0 + rng.uniform(0, 3), lane_y, 0.0, rng.uniform(4, 9))
    return Scenario(p, agents, n_mcts, base.lanes, base.road, rules, state)


def test_agents_leaving_the_map(em, orc):
    """remove_agents_out_of_map: vanished agents, their out-of-map reward, no ego -> terminal."""
    prim = [(0.0, 0.0), (0.0, 0.6), (2.0, -0.4)]
    scn = scenarios.straight_road_test_world(n_others=3, multi_agent=True, primitives=prim, rules=[
        RuleMonitor.MakeRule("G !collision_corridor", -7.0, 0), RuleMonitor.MakeRule("G !collision_ego", -3.0, 1)])
    scn.params.REMOVE_AGENTS_OUT_OF_MAP = True
    scn.params.REWARD_VECTOR_SIZE = 2
    scn.initial_state[1] = (30.0, -5.25, 0.0, 5.0)
    cfg = oracle_config(scn)
    rng = np.random.default_rng(8)
    gone = 0
    for b in range(64):
        acts = rng.integers(0, 3, size=(18, 4)).astype(np.uint8)
        o = orc.replay(cfg, scn.initial_state, FL(4), 20, acts)
        tr = em.replay(scn, scn.initial_state, FL(4), 20, acts)
        compare_trace(tr, 0, o, scn)
        gone += int((o["flags"][:o["steps"]] == 0).any())
    assert gone > 10
    ref = orc.rollout_detail(cfg, scn.initial_state, FL(4), 20, 3, 0, 600, 30, 0.9)
    out = em.rollout_detail(scn, scn.initial_state, FL(4), 20, 3, 0, 600, 30, 0.9)
    compare_rollouts(out, ref, "vanish")
    assert (ref["flags"] == 0).any()


def test_evaluate_rules(em, orc):
    scn = scenarios.merge_scenario(seed=4)
    cfg = oracle_config(scn)
    rng = np.random.default_rng(1)
    hits = 0
    for b in range(64):
        st = scn.initial_state.copy()
        st[:, 0] += rng.uniform(-12, 30, size=4)
        fl = FL(4)
        fl[rng.random(4) < 0.2] = ocw.PRESENT
        q = rng.integers(0, 3, size=(4, 4)).astype(np.uint8)
        q[:, 0] = 0
        want = orc.evaluate_rules(cfg, st, fl, 20, q)
        got = em.evaluate_rules(scn, st, fl, 20, q)
        for k in ("q", "viol", "labels"):
            assert np.array_equal(got[k], want[k]), (b, k)
        assert_close(got["rewards"], want["rewards"], b)
        hits += int(want["viol"].sum() > 0)
    assert hits > 0


def test_terminal_leaf_returns_discounted_terminal_reward(em, orc):
    scn = scenarios.merge_scenario()
    cfg = oracle_config(scn)
    ref = orc.rollout_detail(cfg, scn.initial_state, FL(4), 0, 1, 0, 4, 30, 0.9)   # horizon 0: terminal
    out = em.rollout_detail(scn, scn.initial_state, FL(4), 0, 1, 0, 4, 30, 0.9, leaf_terminal=True)
    compare_rollouts(out, ref, "terminal leaf")
    assert out["agent_steps"] == 0
