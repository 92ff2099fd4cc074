"""GPU parity tests of the GridWorld path, all through the C ABI (libbrm.so):
CUDA kernels vs the CPU oracle (and the reference's own code where oracle/_ref travelled),
bit-exact: integer states, label truth values, automaton states, violation counts AND the
double-precision rewards/returns."""
import os

import numpy as np
import pytest

import planner_rules_mcts_b200 as brm
from oracle import gw as ogw
from tests.golden.make_gw_golden import scenario
from tests.gw_common import (oracle_config_from_engine, oracle_labels_to_words, random_states)

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "gw_golden.npz")
HAVE_REF = os.path.exists(os.path.join(ogw.HERE, "_ref", "libgw_ref.so"))


def make_engine(A=3, terminal_depth=8, on_violation="reset", appendix_b_tables=False, **kw):
    p = brm.GridWorldStateParameter(num_other_agents=A - 1, terminal_depth_=terminal_depth, **kw)
    rules = brm.make_default_rules(on_violation)
    if appendix_b_tables:
        # the golden traces number the ZIP monitor's states as SURVEY.md Appendix B does; the
        # compiler's breadth-first numbering is an isomorphic relabelling (tests/test_ltl.py)
        zip_aps = ["in_direct_front_x#0", "merged_e", "merged_x#0"]
        rules = [brm.RuleMonitor(brm.ltl.RuleTable("G !collision", ["collision"],
                                                   ogw.G_NOT_COLLISION["trans"], [1]), -1.0, 0, on_violation),
                 brm.RuleMonitor(brm.ltl.RuleTable(brm.ZIP_FORMULA, zip_aps, ogw.ZIP["trans"],
                                                   ogw.ZIP["accepting"]), -1.0, 1, on_violation)]
    return brm.GridWorldEngine(p, rules, device=0)


def check_trace(tr, b, o, A):
    k = o["steps"]
    assert tr["steps_out"][b] == k
    assert np.array_equal(tr["agents_out"][b, :k], o["agents"][:k])
    term_bits = (o["term"][:k].astype(np.uint32) << np.arange(A, dtype=np.uint32)).sum(1)
    assert np.array_equal(tr["term_out"][b, :k], term_bits.astype(np.uint8))
    assert np.array_equal(tr["q_out"][b, :k], o["q"][:k])
    assert np.array_equal(tr["viol_out"][b, :k], o["viol"][:k])
    assert np.array_equal(tr["rewards_out"][b, :k], o["rewards"][:k])  # doubles, bit-exact
    assert np.array_equal(tr["labels_out"][b, :k], oracle_labels_to_words(o["labels"][:k], A))
    assert np.array_equal(tr["term_reward_out"][b], o["term_reward"])


def test_replay_matches_golden_from_reference(built):
    g = np.load(GOLDEN)
    for name in ["base3_reset", "base3_stay", "merge4_reset"]:
        cfg, A = scenario(name)
        eng = make_engine(A, cfg.terminal_depth, "reset" if name.endswith("reset") else "stay",
                          appendix_b_tables=True)
        ag, term, acts = g[name + "/agents0"], g[name + "/term0"], g[name + "/actions"]
        tr = eng.replay(eng.make_batch(ag, term), acts)
        for b in range(ag.shape[0]):
            o = {k: g[name + "/" + k][b] for k in
                 ["agents", "term", "q", "viol", "rewards", "labels", "term_reward", "steps"]}
            o["steps"] = int(o["steps"])
            check_trace(tr, b, o, A)


@pytest.mark.parametrize("A,depth,policy", [(1, 6, "reset"), (2, 9, "reset"), (3, 12, "stay"),
                                            (3, 8, "reset"), (4, 10, "reset"), (5, 8, "stay"),
                                            (6, 7, "reset"), (8, 6, "reset")])
def test_replay_matches_oracles_live(built, A, depth, policy):
    eng = make_engine(A, depth, policy)
    cfg = oracle_config_from_engine(eng)
    oracles = [ogw.GwOracle("port")] + ([ogw.GwOracle("ref")] if HAVE_REF else [])
    rng = np.random.default_rng(10 * A + depth)
    B, T = 200, depth + 3
    agents, term = random_states(rng, B, A, p_terminal=0.15)
    acts = rng.integers(0, 3, size=(B, T, A)).astype(np.uint8)
    tr = eng.replay(eng.make_batch(agents, term), acts)
    for orc in oracles:
        for b in range(B):
            check_trace(tr, b, orc.replay(cfg, agents[b], term[b], 0, acts[b]), A)


def test_execute_batch_matches_replay_and_oracle(built):
    """K2: batched Execute + GetNumActions/IsTerminal/GetTerminalReward vs the oracle."""
    A = 3
    eng = make_engine(A, 10)
    cfg = oracle_config_from_engine(eng)
    port = ogw.GwOracle("port")
    rng = np.random.default_rng(5)
    n = 1000
    agents, term = random_states(rng, n, A, p_terminal=0.2)
    batch = eng.make_batch(agents, term)
    acts = rng.integers(0, 3, size=(A, n)).astype(np.uint8)
    nb, rew = eng.execute(batch, acts)
    na = eng.get_num_actions(nb)
    it = eng.is_terminal(nb)
    trw = eng.terminal_reward(nb)
    for i in range(n):
        o = port.replay(cfg, agents[i], term[i], 0, acts[:, i][None])
        assert o["steps"] == 1
        assert np.array_equal(nb.agents[:, i], o["agents"][0])
        assert np.array_equal(nb.q[:, :, i], o["q"][0])
        assert np.array_equal(nb.viol[:, :, i], o["viol"][0])
        assert np.array_equal(rew[:, :, i], o["rewards"][0])
        tbits = int((o["term"][0].astype(np.uint32) << np.arange(A)).sum())
        assert nb.term[i] == tbits and nb.depth[i] == 1
        assert it[i] == bool(o["term"][0, 0])
        assert np.array_equal(na[:, i], np.where(o["term"][0] != 0, 1, 3))
        assert np.array_equal(trw[:, :, i], o["term_reward"])


def test_state_interface_mirror(built):
    """GridWorldState.Execute / GetNumActions / IsTerminal / GetTerminalReward, used the way
    gridworld/test_runner.cpp:28-42 drives the reference."""
    eng = make_engine(3, 8)
    cfg = oracle_config_from_engine(eng)
    port = ogw.GwOracle("port")
    s = brm.GridWorldState(eng)
    assert s.GetAgentIdx() == [0, 1, 2] and not s.IsTerminal() and s.GetNumActions(0) == 3
    rewards = []
    steps = 0
    # ego waits twice so that it reaches the merge point together with agent 2 -> collision
    plan = [[1, 0, 0], [1, 0, 0]] + [[0, 0, 0]] * 18
    agents, term, q = s.GetAgentStates(), np.zeros(3, np.uint8), s.GetRuleStates()
    while not s.IsTerminal() and steps < 20:
        o = port.replay(cfg, agents, term, 0, np.array([plan[steps]], np.uint8), q0=q)
        s = s.Execute(plan[steps], rewards)
        s.ResetDepth()  # closed loop of gridworld/test_runner.cpp:37-38
        agents, term, q = o["agents"][0], o["term"][0], o["q"][0]
        assert np.array_equal(s.GetAgentStates(), agents) and np.array_equal(s.GetRuleStates(), q)
        assert np.array_equal(np.array(rewards), o["rewards"][0])
        steps += 1
    assert s.IsTerminal() and steps == 8 and len(rewards) == 3 and rewards[0].shape == (3,)
    assert rewards[0][0] == -1.0  # `G !collision` violated by the ego, weight -1 at priority 0
    assert s.GetNumActions(0) == 1
    assert s.PrintState().startswith("Ego: x=8")
    tr = s.GetTerminalReward()
    assert len(tr) == 3 and all(np.all(t == 0) for t in tr)  # safety monitors: no final penalty


@pytest.mark.parametrize("A,n,depth", [(3, 4096, 8), (2, 1000, 12), (4, 777, 10), (1, 100, 5)])
def test_rollout_per_rollout_outputs_match_oracle(built, A, n, depth):
    """K1 with Philox actions: every rollout's return vector (double), violation counters,
    final automaton and agent states and length equal the oracle's, bit for bit."""
    eng = make_engine(A, depth)
    cfg = oracle_config_from_engine(eng)
    port = ogw.GwOracle("port")
    rng = np.random.default_rng(A)
    L = 3
    agents, term = random_states(rng, L, A, spread=6)
    leaves = eng.make_batch(agents, term)
    out = eng.rollout(leaves, n, seed=1234, rollout_base=55, max_steps=30, discount_factor=0.95,
                      detail=True)
    steps_total = 0
    for l in range(L):
        ref = port.rollout_detail(cfg, agents[l], term[l], 0, 1234, 55 + l * n, n, 30, 0.95)
        sl = slice(l * n, (l + 1) * n)
        assert np.array_equal(out["ro_returns"][sl], ref["returns"])
        assert np.array_equal(out["ro_viol"][sl], ref["viol"].astype(np.uint8))
        assert np.array_equal(out["ro_q"][sl], ref["q"])
        assert np.array_equal(out["ro_agents"][sl], ref["agents"])
        assert np.array_equal(out["ro_steps"][sl], ref["steps"])
        np.testing.assert_allclose(out["returns_sum"][l], ref["returns"].sum(0), rtol=1e-12, atol=1e-9)
        assert np.array_equal(out["viol_sum"][l], ref["viol"].sum(0).astype(np.uint64))
        agg = port.rollout(cfg, agents[l], term[l], 0, 1234, 55 + l * n, n, 30, 0.95)
        assert out["agent_steps"][l] == agg["agent_steps"]
        steps_total += agg["agent_steps"]
    assert steps_total > 0


def test_rollout_policies_and_determinism(built):
    eng = make_engine(3, 8)
    cfg = oracle_config_from_engine(eng)
    port = ogw.GwOracle("port")
    ag, term = ogw.base_test_env_state()
    leaves = eng.base_test_env_batch(1)
    a = eng.rollout(leaves, 2048, seed=9, max_steps=5, discount_factor=0.9, first_discount_exp=0, detail=True)
    b = eng.rollout(leaves, 2048, seed=9, max_steps=5, discount_factor=0.9, first_discount_exp=0, detail=True)
    for k in a:
        assert np.array_equal(a[k], b[k])  # reproducible, including the summed returns
    ref = port.rollout_detail(cfg, ag, term, 0, 9, 0, 2048, 5, 0.9, first_discount_exp=0)
    assert np.array_equal(a["ro_returns"], ref["returns"]) and a["ro_steps"].max() == 5
    c = eng.rollout(leaves, 2048, seed=10, max_steps=5, discount_factor=0.9, first_discount_exp=0, detail=True)
    assert not np.array_equal(a["ro_returns"], c["ro_returns"])  # the seed matters
    z = eng.rollout(leaves, 64, seed=9, max_steps=0, discount_factor=0.9)
    assert np.all(z["returns_sum"] == 0) and z["agent_steps"][0] == 0


def test_full_size_properties(built):
    """BASELINE-size run (64k rollouts from each of 16 leaves): size-independent properties --
    counts add up, sums are reproducible and split-invariant, device path == host path."""
    import torch
    eng = make_engine(3, 8)
    L, n = 16, 65536
    rng = np.random.default_rng(0)
    agents, term = random_states(rng, L, 3, spread=6)
    leaves = eng.make_batch(agents, term)
    full = eng.rollout(leaves, n, seed=1000)
    assert full["agent_steps"].sum() <= L * n * 8 * 3 and full["agent_steps"].min() > 0
    # split the rollouts of every leaf in two halves with the right rollout ids: sums add up
    # (violation counters exactly, returns to rounding)
    ids = np.arange(L, dtype=np.uint64) * n
    lo = np.zeros_like(full["returns_sum"]); vs = np.zeros_like(full["viol_sum"]); st = 0
    for l in range(L):
        one = leaves.select(np.array([l]))
        for half in range(2):
            o = eng.rollout(one, n // 2, seed=1000, rollout_base=int(ids[l]) + half * (n // 2))
            lo[l] += o["returns_sum"][0]; vs[l] += o["viol_sum"][0]; st += int(o["agent_steps"][0])
    assert np.array_equal(vs, full["viol_sum"]) and st == int(full["agent_steps"].sum())
    np.testing.assert_allclose(lo, full["returns_sum"], rtol=1e-10, atol=1e-8)
    # device-resident path gives the same bits as the host path
    dl = leaves.cuda(0)
    out = dict(returns_sum=torch.zeros((L, 3, 3), dtype=torch.float64, device="cuda"),
               viol_sum=torch.zeros((L, 3, eng.R), dtype=torch.int64, device="cuda"),
               agent_steps=torch.zeros(L, dtype=torch.int64, device="cuda"))
    eng.rollout(dl, n, seed=1000, out=out)
    eng.engine.synchronize()
    assert np.array_equal(out["returns_sum"].cpu().numpy(), full["returns_sum"])
    assert np.array_equal(out["viol_sum"].cpu().numpy().astype(np.uint64), full["viol_sum"])
    # mean ego return is finite and negative (action costs), violations happen but not always
    mean = full["returns_sum"] / n
    assert np.all(np.isfinite(mean)) and np.all(mean[:, :, 2] < 0)
    assert 0 < full["viol_sum"].sum() < L * n * 3 * eng.R * 8


def test_root_stats_reduce(built):
    eng = make_engine(3, 8)
    L, V, A, NA = 27, 3, 3, 3
    rng = np.random.default_rng(1)
    first = np.array([[i // 9, (i // 3) % 3, i % 3] for i in range(L)], np.uint8)
    rs = rng.normal(size=(L, A, V))
    stats = np.zeros((A, NA, 1 + V))
    eng.root_stats_reduce(first, rs, 100, stats)
    for a in range(A):
        for act in range(NA):
            m = first[:, a] == act
            assert stats[a, act, 0] == 100 * m.sum()
            np.testing.assert_allclose(stats[a, act, 1:], rs[m, a].sum(0), rtol=1e-12)


def test_errors(built):
    eng = make_engine(3, 8)
    leaves = eng.base_test_env_batch(1)
    with pytest.raises(brm.BrmError):
        eng.rollout(leaves, 0)
    with pytest.raises(brm.BrmError):
        eng.rollout(leaves, 8, max_steps=1000)
    with pytest.raises(brm.BrmError):
        brm.GridWorldEngine(brm.GridWorldStateParameter(num_other_agents=20), device=0)
    # an empty batch is refused with an error, not a crash
    with pytest.raises(brm.BrmError, match="empty batch"):
        eng.is_terminal(brm.GwBatch.zeros(eng.A, eng.R, 0))
    # Execute must not be given the same arrays for input and output
    import ctypes as C
    from planner_rules_mcts_b200 import _abi
    b = eng.base_test_env_batch(4)
    s = b.to_abi()
    rew = np.zeros((eng.A, eng.V, 4))
    acts = np.zeros((eng.A, 4), np.uint8)
    rc = eng.lib.brm_gw_execute(eng.engine.handle, C.byref(s), acts.ctypes.data, C.byref(s), rew.ctypes.data)
    assert rc == _abi.ERR_INVALID and b"must not share" in eng.lib.brm_last_error()


def test_coop_factor_mixes_agent_returns(built):
    eng = brm.GridWorldEngine()
    leaves = eng.base_test_env_batch(7)
    own = eng.rollout(leaves, 100, seed=3)
    mix = eng.rollout(leaves, 100, seed=3, coop_factor=0.5)
    s = own["returns_sum"]
    np.testing.assert_allclose(mix["returns_sum"], s + 0.5 * (s.sum(axis=1, keepdims=True) - s), rtol=1e-12, atol=1e-9)
    assert (mix["viol_sum"] == own["viol_sum"]).all()


def test_rollout_with_root_stats_in_one_call(built):
    eng = brm.GridWorldEngine()
    rng = np.random.default_rng(4)
    L, n = 17, 64
    leaves = eng.base_test_env_batch(L)
    first = rng.integers(0, 3, size=(L, eng.A)).astype(np.uint8)
    ref = eng.rollout(leaves, n, seed=2)
    want = np.zeros((eng.A, 3, 1 + eng.V))
    eng.root_stats_reduce(first, ref["returns_sum"], n, want)
    got = np.zeros((eng.A, 3, 1 + eng.V))
    out = eng.rollout(leaves, n, seed=2, first_action=first, stats=got)
    assert np.array_equal(got, want) and np.array_equal(out["returns_sum"], ref["returns_sum"])
