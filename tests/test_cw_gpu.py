"""GPU parity tests of the continuous-world path (MvmctsState on the device), all through
the C ABI: CUDA kernels vs the FP64 CPU oracle on replayed action sequences and on
Philox-driven rollouts.  Policy (tests/cw_common.py): labels, automaton states, violation
counts, flags, terminal flags, lane indices and step counts EQUAL on every trace and every
rollout (no exclusions); kinematics / rewards / returns within 1e-9 relative."""
import numpy as np
import pytest

import planner_rules_mcts_b200 as brm
from oracle import cw as ocw
from planner_rules_mcts_b200 import scenarios
from planner_rules_mcts_b200.ltl import RuleMonitor
from tests.cw_common import assert_close, close, compare_rollouts, compare_trace, oracle_config

pytestmark = pytest.mark.gpu

FL = lambda n: np.full(n, ocw.PRESENT | ocw.VALID, np.uint8)


@pytest.fixture(scope="module")
def orc(built):
    return ocw.CwOracle()


def replay_both(eng, orc, scn, acts, states=None, flags=None, horizon=None):
    B = acts.shape[0]
    A = len(scn.agents)
    states = np.broadcast_to(scn.initial_state, (B, A, 4)) if states is None else states
    flags = np.broadcast_to(FL(A), (B, A)) if flags is None else flags
    horizon = scn.params.HORIZON if horizon is None else horizon
    batch = eng.make_batch(states, flags=flags, horizon=horizon)
    tr = eng.replay(batch, acts)
    cfg = oracle_config(scn)
    q0 = eng.initial_q() if scn.agent_rules else None   # rules that do not apply sit in their disabled state
    outs = [orc.replay(cfg, states[b], flags[b], horizon, acts[b], q0=q0) for b in range(B)]
    return tr, outs, batch


def test_reference_pins_single_agent(built, orc):
    """test/single_agent_test.cpp:136-173 through the StateInterface mirror on the GPU."""
    scn = scenarios.straight_road_test_world()
    eng = brm.RulesMctsEngine([scn], device=0)
    s0 = brm.MvmctsState(eng)
    assert not s0.IsTerminal() and s0.GetNumActions(0) == 3 and s0.GetAgentIdx() == [0]
    rewards = []
    s1 = s0.Execute([0], rewards)
    total = rewards[0] + s1.GetTerminalReward()[0]
    assert not s1.IsTerminal() and abs(total[0]) < 1e-5
    s2 = s1.Execute([1], rewards)
    total = total + rewards[0] + s2.GetTerminalReward()[0]
    assert s2.IsTerminal() and abs(total[0] + 800.0) < 1e-5
    s3 = s0.Execute([2], rewards)
    assert s3.IsTerminal() and abs(rewards[0][0] + s3.GetTerminalReward()[0][0] + 1000.0) < 1e-5
    assert s3.GetNumActions(0) == 1
    with pytest.raises(brm.BrmError) as ei:  # BARK_EXPECT_TRUE(!IsTerminal()), src/rules_mcts_state.cpp:59
        s3.Execute([0], rewards)
    assert ei.value.code == brm._abi.ERR_TERMINAL
    s = s0.Execute([0], rewards)
    for _ in range(200):
        if s.IsTerminal():
            break
        s = s.Execute([0], rewards)
    assert s.IsTerminal() and 20.0 <= s.GetAgentStates()[0, 0] - 1.0 and s.GetAgentStates()[0, 0] + 3.0 <= 27.0


def test_reference_pins_multi_agent(built, orc):
    """test/multi_agent_test.cpp:115-139: 3 actions before, 1 after the collision, frozen positions."""
    scn = scenarios.straight_road_test_world(n_others=2, ego_velocity=0.0, velocity_difference=-3.0,
                                             multi_agent=True, primitives=[(0.0, 0.0), (50.0, 1.0), (4.0, 0.0)],
                                             rules=[RuleMonitor.MakeRule("G !collision_ego", -1000.0, 0)])
    scn.initial_state[1] = (10.0, -1.75, 0.0, 3.0)
    scn.initial_state[2] = (14.1, -1.75, 0.0, 3.0)
    eng = brm.RulesMctsEngine([scn], device=0)
    s = brm.MvmctsState(eng)
    assert s.GetNumActions(2) == 3 and s.GetNumActions(1) == 3
    rewards = []
    s2 = s.Execute([0, 2, 0], rewards)
    assert s2.GetNumActions(1) == 1 and s2.GetNumActions(2) == 1 and s2.GetNumActions(0) == 3
    assert rewards[1][0] == -1000.0 and rewards[2][0] == -1000.0 and rewards[0][0] == 0.0
    s3 = s2.Execute([0, 0, 0], rewards)
    assert s3.GetNumActions(1) == 1 and s3.GetNumActions(2) == 1
    assert np.array_equal(s3.GetAgentStates()[1:], s2.GetAgentStates()[1:])
    assert rewards[1][0] == 0.0 and rewards[2][0] == 0.0  # INVALID agents are not evaluated (:83)


@pytest.mark.parametrize("name", ["highway", "merge", "intersection", "test_world"])
def test_replay_matches_oracle(built, orc, name):
    scn = {"highway": scenarios.highway_scenario, "merge": scenarios.merge_scenario,
           "intersection": scenarios.intersection_scenario,
           "test_world": lambda: scenarios.straight_road_test_world(
               n_others=2, multi_agent=True, primitives=[(0.0, 0.0), (3.0, 0.05), (-2.0, -0.05), (4.0, 0.0)])}[name]()
    p = scn.params
    p.ACCELERATION_WEIGHT, p.RADIAL_ACCELERATION_WEIGHT, p.LANE_CENTER_WEIGHT = -0.5, -1.0, -2.0
    p.GOAL_REWARD, p.COLLISION_WEIGHT = 100.0, -50.0
    eng = brm.RulesMctsEngine([scn], device=0)
    rng = np.random.default_rng(11)
    B, T = 96, 22
    n_act = len(scn.agents[0].primitives)
    acts = rng.integers(0, n_act, size=(B, T, scn.n_mcts_agents)).astype(np.uint8)
    tr, outs, _ = replay_both(eng, orc, scn, acts)
    total = sum(compare_trace(tr, b, o, scn) for b, o in enumerate(outs))
    assert total > 3 * B  # traces are not trivially short
    # none of the oracle's threshold comparisons was decided by rounding noise
    assert min(o["margin"][:o["steps"]].min() for o in outs if o["steps"]) > 1e-9


def test_replay_perturbed_states_and_flags(built, orc):
    """Random start states (incl. collided / absent agents, partially advanced automata)."""
    scn = scenarios.merge_scenario(seed=5)
    eng = brm.RulesMctsEngine([scn], device=0)
    rng = np.random.default_rng(2)
    B, T, A = 128, 12, 4
    states = np.broadcast_to(scn.initial_state, (B, A, 4)).copy()
    states[:, :, 0] += rng.uniform(-15, 25, size=(B, A))
    states[:, :, 3] += rng.uniform(-4, 4, size=(B, A))
    flags = np.broadcast_to(FL(A), (B, A)).copy()
    flags[rng.random((B, A)) < 0.1] = ocw.PRESENT          # already collided (INVALID)
    flags[:, 1:][rng.random((B, A - 1)) < 0.05] = 0         # absent
    flags[:, 0] |= ocw.PRESENT
    acts = rng.integers(0, 3, size=(B, T, A)).astype(np.uint8)
    tr, outs, batch = replay_both(eng, orc, scn, acts, states, flags, 9)
    cfg = oracle_config(scn)
    n_term0 = 0
    for b, o in enumerate(outs):
        t0 = orc.check_terminal(cfg, states[b], flags[b], 9)
        assert bool(batch.terminal[b]) == t0
        n_term0 += t0
        compare_trace(tr, b, o, scn)
    assert 0 < n_term0 < B


def test_execute_batch_matches_replay(built, orc):
    """K4: batched Execute chained 3 times == replay; GetNumActions / GetTerminalReward."""
    scn = scenarios.merge_scenario(seed=9)
    eng = brm.RulesMctsEngine([scn], device=0)
    rng = np.random.default_rng(4)
    n, A = 300, 4
    states = np.broadcast_to(scn.initial_state, (n, A, 4)).copy()
    states[:, :, 0] += rng.uniform(-10, 10, size=(n, A))
    batch = eng.make_batch(states)
    keep = np.flatnonzero(batch.terminal == 0)
    batch = batch.select(keep)
    acts = rng.integers(0, 3, size=(len(keep), 3, A)).astype(np.uint8)
    tr = eng.replay(batch, acts)
    cur = batch
    alive = np.arange(len(keep))
    for t in range(3):
        nb, rew = eng.execute(cur, acts[alive, t].T.copy())
        # the replay kernel carries the Kahan compensation of the integrator across steps, a
        # chain of single Executes restarts it every step: identical at t = 0, ~1 ulp later
        # a chain of single Executes carries the same FP64 state as the replay kernel: same bits
        assert np.array_equal(nb.agents.transpose(1, 0, 2), tr["agents_out"][alive, t])
        assert np.array_equal(nb.flags.T, tr["flags_out"][alive, t])
        assert np.array_equal(nb.q.transpose(2, 0, 1), tr["q_out"][alive, t])
        assert np.array_equal(nb.viol.transpose(2, 0, 1), tr["viol_out"][alive, t])
        assert np.array_equal(rew.transpose(2, 0, 1), tr["rewards_out"][alive, t])
        assert np.array_equal(nb.terminal, tr["terminal_out"][alive, t])
        assert np.all(nb.horizon == scn.params.HORIZON - t - 1)
        na = eng.get_num_actions(nb)
        movable = (nb.flags & 3) == 3
        assert np.array_equal(na, np.where(movable, 3, 1))
        assert eng.terminal_reward(nb).shape == (A, 3, nb.n)
        sel = np.flatnonzero(nb.terminal == 0)
        cur, alive = nb.select(sel), alive[sel]
        if len(sel) == 0:
            break
    with pytest.raises(brm.BrmError):
        term = nb.select(np.flatnonzero(nb.terminal != 0)[:1]) if (nb.terminal != 0).any() else None
        if term is None:
            raise brm.BrmError(-4, "no terminal state to test")
        eng.execute(term, np.zeros((A, 1), np.uint8))


@pytest.mark.parametrize("name,n", [("merge", 4096), ("highway", 2048), ("intersection", 300)])
def test_rollout_per_rollout_outputs_match_oracle(built, orc, name, n):
    """K3 with Philox actions vs the oracle's rollouts, rollout by rollout."""
    scn = {"highway": scenarios.highway_scenario, "merge": scenarios.merge_scenario,
           "intersection": scenarios.intersection_scenario}[name](seed=21)
    eng = brm.RulesMctsEngine([scn], device=0)
    cfg = oracle_config(scn)
    A = len(scn.agents)
    leaves = eng.initial_batch()
    out = eng.rollout(leaves, n, seed=77, rollout_base=1000, max_steps=30, detail=True)
    ref = orc.rollout_detail(cfg, scn.initial_state, FL(A), scn.params.HORIZON, 77, 1000, n, 30,
                             scn.params.DISCOUNT_FACTOR)
    compare_rollouts(out, ref, name)       # every rollout: steps, violations, automaton states, flags EQUAL
    assert ref["margin"].min() > 1e-9       # ... and none of them by luck: no comparison sat on its threshold
    # aggregates
    agg = orc.rollout(cfg, scn.initial_state, FL(A), scn.params.HORIZON, 77, 1000, n, 30,
                      scn.params.DISCOUNT_FACTOR, n_threads=4)
    assert int(out["agent_steps"][0]) == agg["agent_steps"]
    assert np.array_equal(out["viol_sum"][0], agg["viol_sum"])
    # per-leaf sums are accumulated in 2^-24 fixed point (order-independent integer atomics)
    np.testing.assert_allclose(out["returns_sum"][0], out["ro_returns"].sum(0), rtol=0, atol=n * 2.0 ** -24)
    np.testing.assert_allclose(out["returns_sum"][0], agg["returns_sum"], rtol=1e-9, atol=n * 2.0 ** -24)
    assert np.array_equal(out["viol_sum"][0], out["ro_viol"].astype(np.uint64).sum(0))


def test_multi_scenario_sweep_and_determinism(built, orc):
    """Config W shape: several scenarios in one rollout call, 64 rollouts each."""
    scns = [scenarios.intersection_scenario(seed=1000, scenario_id=k) for k in range(6)]
    eng = brm.RulesMctsEngine(scns, device=0)
    leaves = eng.initial_batch()
    assert leaves.n == 6 and list(leaves.scenario) == list(range(6))
    a = eng.rollout(leaves, 64, seed=3, detail=True)
    b = eng.rollout(leaves, 64, seed=3, detail=True)
    for k in a:
        assert np.array_equal(a[k], b[k])
    for k, scn in enumerate(scns):
        cfg = oracle_config(scn)
        ref = orc.rollout_detail(cfg, scn.initial_state, FL(10), 20, 3, k * 64, 64, 30, 0.9)
        sl = slice(k * 64, (k + 1) * 64)
        compare_rollouts({k2: v[sl] for k2, v in a.items() if k2.startswith("ro_")}, ref, "scenario %d" % k)
    assert len({float(x) for x in a["returns_sum"][:, 0, 3]}) == 6  # scenarios really differ


def test_few_leaves_shared_by_many_ctas(built, orc):
    """Several scenarios, few leaves, many rollouts per leaf: the CTAs that find no fresh leaf join
    the leaves that still have unclaimed rollouts.  Every rollout runs exactly once, whoever ran it."""
    scns = [scenarios.intersection_scenario(seed=1000, scenario_id=k) for k in range(3)]
    eng = brm.RulesMctsEngine(scns, device=0)
    leaves = eng.initial_batch()
    n = 3000
    a = eng.rollout(leaves, n, seed=5, detail=True)
    b = eng.rollout(leaves, n, seed=5, detail=True)
    for k in a:
        assert np.array_equal(a[k], b[k]), k
    assert (a["ro_steps"] > 0).all()
    for k, scn in enumerate(scns):
        sl = slice(k * n, (k + 1) * n)
        assert np.array_equal(a["viol_sum"][k], a["ro_viol"][sl].astype(np.uint64).sum(0))
        fx = np.rint(a["ro_returns"][sl] * 2.0 ** 24).astype(np.int64).sum(0)
        assert np.array_equal(a["returns_sum"][k], fx.astype(np.float64) / 2.0 ** 24)
        cfg = oracle_config(scn)
        for lo in (0, n - 40):   # the first and the last rollouts of the leaf against the oracle
            ref = orc.rollout_detail(cfg, scn.initial_state, FL(10), 20, 5, k * n + lo, 40, 30, 0.9)
            part = slice(k * n + lo, k * n + lo + 40)
            compare_rollouts({k2: v[part] for k2, v in a.items() if k2.startswith("ro_")}, ref,
                             "scenario %d rollouts %d.." % (k, lo))


def test_full_size_properties(built):
    """BASELINE config M size (262144 rollouts/decision): split invariance, device path ==
    host path, counts add up."""
    import torch
    scn = scenarios.merge_scenario(seed=1000)
    eng = brm.RulesMctsEngine([scn], device=0)
    L, n = 64, 4096
    rng = np.random.default_rng(0)
    states = np.broadcast_to(scn.initial_state, (L, 4, 4)).copy()
    states[:, :, 0] += rng.uniform(-5, 5, size=(L, 4))
    leaves = eng.make_batch(states)
    full = eng.rollout(leaves, n, seed=1000, max_steps=20)
    assert full["agent_steps"].sum() <= L * n * 20 * 4
    half = [eng.rollout(leaves.select([l]), n // 2, seed=1000, max_steps=20, rollout_base=l * n + h * (n // 2))
            for l in range(4) for h in range(2)]
    for l in range(4):
        a, b = half[2 * l], half[2 * l + 1]
        assert np.array_equal(a["viol_sum"][0] + b["viol_sum"][0], full["viol_sum"][l])
        assert int(a["agent_steps"][0] + b["agent_steps"][0]) == int(full["agent_steps"][l])
        # per-leaf sums are accumulated in 2^-24 fixed point (order-independent integer atomics): exact
        assert np.array_equal(a["returns_sum"][0] + b["returns_sum"][0], full["returns_sum"][l])
    dl = leaves.cuda(0)
    out = dict(returns_sum=torch.zeros((L, 4, 3), dtype=torch.float64, device="cuda"),
               viol_sum=torch.zeros((L, 4, eng.R), dtype=torch.int64, device="cuda"),
               agent_steps=torch.zeros(L, dtype=torch.int64, device="cuda"))
    eng.rollout(dl, n, seed=1000, max_steps=20, out=out)
    eng.engine.synchronize()
    assert np.array_equal(out["returns_sum"].cpu().numpy(), full["returns_sum"])
    assert np.array_equal(out["viol_sum"].cpu().numpy().astype(np.uint64), full["viol_sum"])
    assert np.array_equal(out["agent_steps"].cpu().numpy().astype(np.uint64), full["agent_steps"])


def test_errors(built):
    scn = scenarios.merge_scenario()
    eng = brm.RulesMctsEngine([scn], device=0)
    with pytest.raises(brm.BrmError):
        eng.rollout(eng.initial_batch(), 0)
    bad = scenarios.merge_scenario()
    bad.rules = [RuleMonitor.MakeRule("G !collision_ego", -1.0, 5)]  # priority >= V
    with pytest.raises(brm.BrmError):
        brm.RulesMctsEngine([bad], device=0)
    other = scenarios.highway_scenario()
    with pytest.raises(brm.BrmError):
        brm.RulesMctsEngine([scn, other], device=0)  # shapes differ
    # an empty batch is refused with an error, not a crash
    empty = brm.CwBatch.zeros(eng.A, eng.M, eng.R, 0)
    with pytest.raises(brm.BrmError, match="empty batch"):
        eng.get_num_actions(empty)
    with pytest.raises(brm.BrmError, match="empty batch"):
        eng.rollout(empty, 4)
    # more rollouts in one call than the 31-bit rollout index holds
    with pytest.raises(brm.BrmError, match="rollouts in one call"):
        eng.rollout(eng.initial_batch(), 2**31 - 1)


def test_replay_matches_committed_golden_traces(built):
    """CUDA kernels vs tests/golden/cw_golden.npz (FP64 oracle traces frozen in the repo)."""
    import os
    from tests.golden.make_cw_golden import SCENARIOS
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cw_golden.npz"))
    for name, make in SCENARIOS.items():
        scn = make()
        eng = brm.RulesMctsEngine([scn], device=0)
        acts = g[name + "/actions"]
        B = acts.shape[0]
        batch = eng.make_batch(np.broadcast_to(scn.initial_state, (B,) + scn.initial_state.shape))
        tr = eng.replay(batch, acts)
        compared = 0
        for b in range(B):
            o = {k: g[name + "/" + k][b] for k in
                 ["states", "flags", "terminal", "q", "viol", "rewards", "labels", "frenet", "margin", "term_reward"]}
            o["steps"] = int(g[name + "/steps"][b])
            compared += compare_trace(tr, b, o, scn)
        assert compared == int(g[name + "/steps"].sum())


@pytest.mark.parametrize("n_agents,n_mcts,n", [(1, 1, 333), (2, 1, 100), (7, 7, 65), (16, 16, 50), (16, 3, 31),
                                               (16, 1, 40), (5, 2, 77)])
def test_edge_shapes_rollouts_match_oracle(built, orc, n_agents, n_mcts, n):
    """Extreme shapes: one agent, the 16-agent maximum (2 rollouts per warp), rollout counts that
    do not divide the lane groups, a mix of controlled and passive agents; many rule slots."""
    rng = np.random.default_rng(n_agents * 17 + n_mcts)
    base = scenarios.straight_road_test_world(n_others=0)
    p = brm.RulesMctsStateParameters(REWARD_VECTOR_SIZE=3, HORIZON=12)
    prim = [(0.0, 0.0), (2.0, 0.0), (-3.0, 0.0), (0.0, 0.03)]
    agents = []
    for i in range(n_agents):
        agents.append(brm.AgentModel(primitives=prim if i < n_mcts else [(0.5, 0.0)], front=3.0, rear=1.0,
                                     half_width=0.9, goal=(150.0, 400.0, -7.0, 0.0) if i == 0 else None))
    rules = [RuleMonitor.MakeRule("G !collision_ego", -100.0, 0),
             RuleMonitor.MakeRule("G safe_distance_front", -1.0, 1),
             RuleMonitor.MakeRule("G !(near_x#0 & in_direct_front_x#0)", -1.0, 1)]
    state = np.zeros((n_agents, 4), np.float32)
    for i in range(n_agents):
        lane_y = -1.75 if i % 2 == 0 else -5.25
        state[i] = ((i // 2) * 14.0 + rng.uniform(0, 3), lane_y, 0.0, rng.uniform(4, 9))
    scn = brm.Scenario(p, agents, n_mcts, base.lanes, base.road, rules, state)
    eng = brm.RulesMctsEngine([scn], device=0)
    assert eng.R == 2 + (n_agents - 1)
    cfg = oracle_config(scn)
    out = eng.rollout(eng.initial_batch(), n, seed=5, max_steps=30, detail=True)
    ref = orc.rollout_detail(cfg, scn.initial_state, FL(n_agents), 12, 5, 0, n, 30, p.DISCOUNT_FACTOR)
    compare_rollouts(out, ref, (n_agents, n_mcts))
    assert ref["margin"].min() > 1e-9
    assert int(out["agent_steps"][0]) > 0
    # leaves that are already terminal produce only the (discounted) terminal reward
    term = eng.make_batch(state[None], horizon=0)
    assert term.terminal[0] == 1
    z = eng.rollout(term, 8, seed=1)
    assert z["agent_steps"][0] == 0 and np.all(z["viol_sum"] == 0)


def test_coop_factor_mixes_agent_returns(built, orc):
    """COOP_FACTOR (src/util.cpp:28-29): returns_sum_i = own_i + c * sum of the others' (unpinned:
    0 in every test of the reference); per-rollout returns stay unmixed.  Against the oracle's own
    mixing (oracle/cw_oracle.cpp cwo_coop_mix / cwo_rollout), not against the GPU's sums."""
    scn = scenarios.merge_scenario()
    eng = brm.RulesMctsEngine([scn])
    cfg = oracle_config(scn)
    leaves = eng.initial_batch([0]).select(np.zeros(5, np.int64))
    n = 96
    own = eng.rollout(leaves, n, seed=7, max_steps=12, detail=True)
    mix = eng.rollout(leaves, n, seed=7, max_steps=12, detail=True, coop_factor=0.25)
    for l in range(5):
        ref = orc.rollout(cfg, scn.initial_state, FL(4), scn.params.HORIZON, 7, l * n, n, 12,
                          scn.params.DISCOUNT_FACTOR, coop_factor=0.25)
        np.testing.assert_allclose(mix["returns_sum"][l], ref["returns_sum"], rtol=1e-9, atol=4 * n * 2.0 ** -24)
        np.testing.assert_allclose(mix["returns_sum"][l], orc.coop_mix(own["returns_sum"][l], 0.25), rtol=1e-12)
        assert not np.allclose(mix["returns_sum"][l], own["returns_sum"][l])
    assert (mix["ro_returns"] == own["ro_returns"]).all() and (mix["viol_sum"] == own["viol_sum"]).all()


def test_rollout_with_root_stats_in_one_call(built):
    """brm_cw_rollout_root_stats == brm_cw_rollout followed by brm_root_stats_reduce, host and device."""
    import torch
    scn = scenarios.merge_scenario()
    eng = brm.RulesMctsEngine([scn])
    rng = np.random.default_rng(3)
    L, n, n_act = 23, 40, 3
    leaves = eng.initial_batch([0]).select(np.zeros(L, np.int64))
    first = rng.integers(0, n_act, size=(L, eng.M)).astype(np.uint8)
    ref = eng.rollout(leaves, n, seed=11, max_steps=12)
    want = np.ones((eng.M, n_act, 1 + eng.V))          # accumulated INTO: start from non-zero statistics
    eng.root_stats_reduce(first, ref["returns_sum"], n, want, n_act)
    got = np.ones((eng.M, n_act, 1 + eng.V))
    out = eng.rollout(leaves, n, seed=11, max_steps=12, first_action=first, stats=got)
    assert np.array_equal(got, want) and np.array_equal(out["returns_sum"], ref["returns_sum"])
    assert got[:, :, 0].sum() == eng.M * (n_act + L * n)
    # device batches
    d_out = dict(returns_sum=torch.zeros((L, eng.M, eng.V), dtype=torch.float64, device="cuda"))
    d_stats = torch.ones((eng.M, n_act, 1 + eng.V), dtype=torch.float64, device="cuda")
    d_leaves, d_first = leaves.cuda(0), torch.from_numpy(first).cuda()
    torch.cuda.synchronize()
    eng.rollout(d_leaves, n, seed=11, max_steps=12, out=d_out, first_action=d_first, stats=d_stats)
    torch.cuda.synchronize()
    assert np.array_equal(d_stats.cpu().numpy(), want)


def test_agent_rule_masks_match_oracle(built, orc):
    """Rules given to some agents only (BehaviorRulesMcts::AddAgentRules): the slots of rules that do
    not apply to an agent start in the rule's absorbing `disabled_state`, stay there and earn
    nothing -- replay and rollouts against the oracle from the same initial automaton states."""
    scn = scenarios.merge_scenario()
    scn.agent_rules = [[0, 1], [0], [1], []]          # ego: both rules, agent 3: none
    eng = brm.RulesMctsEngine([scn], device=0)
    rng = np.random.default_rng(21)
    B, T = 64, 20
    acts = rng.integers(0, 3, size=(B, T, scn.n_mcts_agents)).astype(np.uint8)
    tr, outs, _ = replay_both(eng, orc, scn, acts)
    for b, o in enumerate(outs):
        compare_trace(tr, b, o, scn)
    q0 = eng.initial_q()
    R0 = 1                                             # slot 0: safe distance, slots 1..3: zipper per other agent
    ran = np.arange(T)[None, :] < tr["steps_out"][:, None]        # steps a replay executed before it ended
    assert (tr["q_out"][:, :, 1, R0:][ran] == q0[1, R0:]).all()   # agent 1 has no zipper rule
    assert (tr["q_out"][:, :, 2, :R0][ran] == q0[2, :R0]).all()   # agent 2 has no safe-distance rule
    assert (tr["q_out"][:, :, 3][ran] == q0[3]).all() and (tr["viol_out"][:, :, 3] == 0).all()
    assert (q0[3] == [r.disabled_state for r in (scn.rules[0], scn.rules[1], scn.rules[1], scn.rules[1])]).all()
    # against the same world with every rule for everybody, the masks do change the outcome
    eng_all = brm.RulesMctsEngine([scenarios.merge_scenario()], device=0)
    tr_all = eng_all.replay(eng_all.make_batch(np.broadcast_to(scn.initial_state, (B, 4, 4))), acts)
    assert (tr_all["viol_out"][:, :, 3].sum() > 0) or (tr_all["q_out"][:, :, 3] != q0[3]).any()
    # rollouts: per-rollout outputs against the oracle
    leaves = eng.initial_batch([0])
    out = eng.rollout(leaves, 48, seed=4, max_steps=12, detail=True)
    ref = orc.rollout_detail(oracle_config(scn), scn.initial_state, FL(4), scn.params.HORIZON, 4, 0, 48, 12,
                             scn.params.DISCOUNT_FACTOR, q0=q0)
    compare_rollouts(out, ref, "agent rules")


def test_gpu_bits_equal_the_host_compiled_work_items(built):
    """tests/cpp/cw_emul.cpp runs the kernels' work-item functions (csrc/cw_core.cuh) on the CPU; the
    library is compiled with --fmad=false and spells its fused multiply-adds out, so the GPU must
    produce the same BITS -- states, returns, everything (a scheduling bug in the pool kernel, a
    stale slot, a lost update would show here even where the oracle comparison has tolerance)."""
    from tests.cw_emul import CwEmul
    em = CwEmul()
    for make, n in ((scenarios.merge_scenario, 1500), (scenarios.highway_scenario, 800),
                    (scenarios.intersection_scenario, 200)):
        scn = make(seed=33)
        eng = brm.RulesMctsEngine([scn], device=0)
        A = len(scn.agents)
        out = eng.rollout(eng.initial_batch(), n, seed=9, rollout_base=50, max_steps=30, detail=True)
        ref = em.rollout_detail(scn, scn.initial_state, FL(A), scn.params.HORIZON, 9, 50, n, 30,
                                scn.params.DISCOUNT_FACTOR)
        for k in ("ro_returns", "ro_viol", "ro_q", "ro_agents", "ro_flags", "ro_steps"):
            assert np.array_equal(out[k], ref[k]), (make.__name__, k)
        assert int(out["agent_steps"][0]) == ref["agent_steps"]


def test_agents_leaving_the_map_on_the_gpu(built, orc):
    """remove_agents_out_of_map (src/rules_mcts_state.cpp:121-125): vanished agents, their
    out-of-map reward, one action afterwards, no ego -> terminal; replay and rollouts vs the oracle."""
    prim = [(0.0, 0.0), (0.0, 0.6), (2.0, -0.4)]
    scn = scenarios.straight_road_test_world(n_others=3, multi_agent=True, primitives=prim, rules=[
        RuleMonitor.MakeRule("G !collision_corridor", -7.0, 0), RuleMonitor.MakeRule("G !collision_ego", -3.0, 1)])
    scn.params.REMOVE_AGENTS_OUT_OF_MAP = True
    scn.params.REWARD_VECTOR_SIZE = 2
    scn.initial_state[1] = (30.0, -5.25, 0.0, 5.0)
    eng = brm.RulesMctsEngine([scn], device=0)
    rng = np.random.default_rng(8)
    acts = rng.integers(0, 3, size=(96, 18, 4)).astype(np.uint8)
    tr, outs, _ = replay_both(eng, orc, scn, acts)
    gone = 0
    for b, o in enumerate(outs):
        compare_trace(tr, b, o, scn)
        gone += int((o["flags"][:o["steps"]] == 0).any())
    assert gone > 10
    out = eng.rollout(eng.initial_batch(), 2000, seed=3, max_steps=30, detail=True)
    ref = orc.rollout_detail(oracle_config(scn), scn.initial_state, FL(4), scn.params.HORIZON, 3, 0, 2000, 30,
                             scn.params.DISCOUNT_FACTOR)
    compare_rollouts(out, ref, "vanish")
    assert (ref["flags"] == 0).any()
    # StateInterface view: a vanished agent has one action, its slot earns OUT_OF_MAP_WEIGHT once
    two = scenarios.straight_road_test_world(n_others=1, multi_agent=True, primitives=[(0.0, 0.0), (0.0, 0.6)],
                                             rules=[RuleMonitor.MakeRule("G !collision_corridor", -7.0, 0)])
    two.params.REMOVE_AGENTS_OUT_OF_MAP = True
    two.initial_state[1] = (30.0, -5.25, 0.0, 5.0)
    s = brm.MvmctsState(brm.RulesMctsEngine([two], device=0))
    rewards = []
    for _ in range(12):
        s = s.Execute([0, 1], rewards)
        if s.GetAgentFlags()[1] == 0:
            break
    assert s.GetAgentFlags()[1] == 0 and rewards[1][0] == -800.0 and s.GetNumActions(1) == 1
    assert not s.IsTerminal() and s.GetMultiAgentRuleState()[1, 0] == 0     # it never saw itself off the road
    s2 = s.Execute([0, 0], rewards)
    assert rewards[1][0] == 0.0 and np.all(s2.GetTerminalReward()[1] == 0.0)


def test_evaluate_rules_matches_oracle(built, orc):
    """brm_cw_evaluate_rules (MvmctsState::EvaluateRules(), :186-197) against the oracle's restatement."""
    scn = scenarios.merge_scenario(seed=4)
    eng = brm.RulesMctsEngine([scn], device=0)
    cfg = oracle_config(scn)
    rng = np.random.default_rng(1)
    n = 200
    states = np.broadcast_to(scn.initial_state, (n, 4, 4)).copy()
    states[:, :, 0] += rng.uniform(-12, 30, size=(n, 4))
    flags = np.broadcast_to(FL(4), (n, 4)).copy()
    flags[rng.random((n, 4)) < 0.2] = ocw.PRESENT
    q = rng.integers(0, 3, size=(n, 4, 4)).astype(np.uint8)
    q[:, :, 0] = 0
    batch = eng.make_batch(states, flags=flags, q=q)
    nb, rew = eng.evaluate_rules(batch)
    hits = 0
    for i in range(n):
        want = orc.evaluate_rules(cfg, states[i], flags[i], scn.params.HORIZON, q[i])
        assert np.array_equal(nb.q[:, :, i], want["q"]), i
        assert np.array_equal(nb.viol[:, :, i], want["viol"]), i
        assert_close(rew[:, :, i], want["rewards"], i)
        hits += int(want["viol"].sum() > 0)
    assert hits > 0 and np.array_equal(nb.agents, batch.agents) and np.array_equal(nb.horizon, batch.horizon)


def test_one_rollout_per_leaf(built, orc):
    """mvmcts::RandomHeuristic runs ONE rollout per expanded node: thousands of different
    leaves with one rollout each in one call, every rollout equal to the oracle's from that leaf."""
    scn = scenarios.merge_scenario(seed=12)
    eng = brm.RulesMctsEngine([scn], device=0)
    cfg = oracle_config(scn)
    rng = np.random.default_rng(6)
    L = 3000
    states = np.broadcast_to(scn.initial_state, (L, 4, 4)).copy()
    states[:, :, 0] += rng.uniform(-6, 6, size=(L, 4))
    states[:, :, 3] += rng.uniform(-2, 2, size=(L, 4))
    leaves = eng.make_batch(states, horizon=15)
    out = eng.rollout(leaves, 1, seed=21, rollout_base=7, max_steps=30, detail=True)
    for l in range(0, L, 7):
        ref = orc.rollout_detail(cfg, states[l], FL(4), 15, 21, 7 + l, 1, 30, scn.params.DISCOUNT_FACTOR)
        if leaves.terminal[l]:
            assert out["ro_steps"][l] == 0
        compare_rollouts({k: v[l:l + 1] for k, v in out.items() if k.startswith("ro_")}, ref, l)
        assert_close(out["returns_sum"][l], ref["returns"][0], l, atol=2.0 ** -23)
