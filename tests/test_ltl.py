"""LTL -> monitor-table compiler (host side of ltl::RuleMonitor::MakeRule)."""
import itertools
import random

import pytest

import planner_rules_mcts_b200 as brm
from oracle import gw as ogw
from planner_rules_mcts_b200 import ltl


def test_zip_formula_matches_appendix_b():
    # hand-derived 4-state monitor of
    # /root/reference/gridworld/base_test_env.cpp:16-19 (validated there by brute force)
    t = brm.compile_rule(brm.ZIP_FORMULA)
    assert t.aps == ["in_direct_front_x#0", "merged_e", "merged_x#0"]
    ref = ltl.RuleTable("zip", t.aps, ogw.ZIP["trans"], ogw.ZIP["accepting"])
    assert t.n_states == 4 and t.isomorphic(ref)
    assert t.is_agent_specific()


def test_g_not_collision():
    # pinned: test/single_agent_test.cpp:143,159 (no violation -> 0, violation -> weight,
    # FinalTransit adds nothing for a safety rule)
    t = brm.compile_rule("G !collision")
    assert t.trans == [[0, ltl.VIOLATION]] and t.accepting == [1]
    assert not t.is_agent_specific()


def test_spot_precedence_u_binds_tighter_than_and():
    a = ltl.parse("a & b U c")
    b = ltl.parse("a & (b U c)")
    assert a == b and a != ltl.parse("(a & b) U c")
    assert ltl.parse("a -> b -> c") == ltl.parse("a -> (b -> c)")
    assert ltl.parse("G a & b") == ltl.parse("(G a) & b")


FORMULAS = ["G !a", "F a", "a U b", "G (a -> X b)", "G (a -> F b)", "(a W b) & G !c",
            "a R b", "(a & !b & (a | c) U b) -> G (b & c -> !a)", "G (a -> (b U c))",
            "X X a", "G (a <-> X !a)", "F (a & X b)", "a M b"]


@pytest.mark.parametrize("formula", FORMULAS)
def test_monitor_agrees_with_direct_finite_trace_semantics(formula):
    t = brm.compile_rule(formula)
    rng = random.Random(hash(formula) & 0xffff)
    aps = t.aps
    for length in range(1, 7):
        for _ in range(60):
            trace = [{a: rng.random() < 0.5 for a in aps} for _ in range(length)]
            # first violation position, if any
            q, bad = t.init, None
            for i, val in enumerate(trace):
                nxt = t.trans[q][t.letter_of(val)]
                if nxt == ltl.VIOLATION:
                    bad = i
                    break
                q = nxt
            if bad is not None:
                # a bad prefix cannot be extended to a satisfying trace
                assert not ltl.satisfies(formula, trace[:bad + 1])
                for ext in range(3):
                    more = trace[:bad + 1] + [{a: rng.random() < 0.5 for a in aps} for _ in range(ext)]
                    assert not ltl.satisfies(formula, more)
            else:
                assert bool(t.accepting[q]) == ltl.satisfies(formula, trace)


def test_exhaustive_short_traces_zip():
    t = brm.compile_rule(brm.ZIP_FORMULA)
    letters = [dict(zip(t.aps, bits)) for bits in itertools.product([False, True], repeat=3)]
    for length in (1, 2, 3):
        for trace in itertools.product(letters, repeat=length):
            viol, pending, _ = t.run(list(trace), reset_on_violation=False)
            sat = ltl.satisfies(brm.ZIP_FORMULA, list(trace))
            assert sat == (viol == 0 and not pending)


def test_rule_packs_into_abi():
    r = brm.RuleMonitor.MakeRule(brm.ZIP_FORMULA, -1.0, 1)
    from planner_rules_mcts_b200.gridworld import GW_LABELS
    abi = r.to_abi(GW_LABELS)
    # 4 automaton states + the absorbing, accepting "rule does not apply" state behind them
    assert abi.n_aps == 3 and abi.n_states == 5 and abi.priority == 1 and abi.weight == -1.0
    assert r.disabled_state == 4 and abi.accepting[4] == 1
    assert all(abi.trans[4 * 8 + l] == 4 for l in range(8))
    assert list(abi.ap_label[:3]) == [3, 1, 2] and list(abi.ap_agent_specific[:3]) == [1, 0, 1]
    with pytest.raises(KeyError):
        brm.RuleMonitor.MakeRule("G !nonsense", -1, 0).to_abi(GW_LABELS)
    with pytest.raises(ValueError):
        brm.compile_rule("a & b & c & d & e")  # more than 4 APs


def test_rejects_garbage():
    for bad in ["a &", "(a", "a b", "G", "a U", "#1"]:
        with pytest.raises(ValueError):
            brm.compile_rule(bad)
