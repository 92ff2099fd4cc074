"""Host-side LTL -> deterministic monitor table compiler.

Replaces what `ltl::RuleMonitor::MakeRule(formula, weight, priority)` does with Spot in
the reference's un-vendored dependency bark-simulator/rule-monitoring @ 187c125a (call
sites /root/reference/gridworld/base_test_env.cpp:101-102,
/root/reference/test/single_agent_test.cpp:62).  Spot is not available, so this module
builds the monitor by formula progression under finite-trace (LTLf) semantics:

  * parse with Spot's operator precedence (unary  !, X, F, G  >  U, R, W, M
    (right-assoc)  >  &  >  |  >  ->, <->  (right-assoc)),
  * states = progressed residual formulas, identified by their truth table over the
    formula's temporal atoms (so Boolean-equivalent residuals merge),
  * a residual equivalent to `false` is a bad prefix: table entry VIOLATION,
  * `accepting[q]` = the residual accepts the empty suffix (G/R true, F/U/X false):
    RuleMonitor::FinalTransit returns 0 there and the rule weight otherwise,
  * Moore minimisation, then breadth-first renumbering from the initial state.

Agent placeholders are written `label#0` (base_test_env.cpp:16-19); a rule that has one
is instantiated once per other agent by the engine.  Semantics of the real library are
not pinned by the reference's tests beyond test/single_agent_test.cpp:143,159
(SURVEY.md Appendix B); the ZIP formula compiles to the table given there.
"""
import re
from collections import deque

VIOLATION = 0xFF
MAX_APS = 4
MAX_STATES = 16

_TOKEN = re.compile(r"\s*(<->|->|&&|\|\||[!&|()]|[A-Za-z_][A-Za-z0-9_]*(?:#\d+)?|[01])")
_UNARY = {"!", "X", "F", "G"}
_BINARY_TEMPORAL = {"U", "R", "W", "M"}
_KEYWORDS = _UNARY | _BINARY_TEMPORAL | {"true", "false", "1", "0"}

TRUE = ("true",)
FALSE = ("false",)


def _tokenize(text):
    pos, out = 0, []
    text = text.strip()
    while pos < len(text):
        m = _TOKEN.match(text, pos)
        if not m:
            raise ValueError("cannot tokenize LTL formula at: %r" % text[pos:])
        out.append(m.group(1))
        pos = m.end()
    return out


def mk_not(f):
    if f == TRUE:
        return FALSE
    if f == FALSE:
        return TRUE
    if f[0] == "not":
        return f[1]
    return ("not", f)


def mk_and(f, g):
    if f == FALSE or g == FALSE:
        return FALSE
    if f == TRUE:
        return g
    if g == TRUE:
        return f
    if f == g:
        return f
    return ("and", f, g)


def mk_or(f, g):
    if f == TRUE or g == TRUE:
        return TRUE
    if f == FALSE:
        return g
    if g == FALSE:
        return f
    if f == g:
        return f
    return ("or", f, g)


class _Parser:
    def __init__(self, tokens):
        self.t = tokens
        self.i = 0

    def peek(self):
        return self.t[self.i] if self.i < len(self.t) else None

    def take(self):
        tok = self.peek()
        self.i += 1
        return tok

    def parse(self):
        f = self.implication()
        if self.peek() is not None:
            raise ValueError("unexpected token %r" % self.peek())
        return f

    def implication(self):
        lhs = self.disjunction()
        tok = self.peek()
        if tok == "->":
            self.take()
            rhs = self.implication()
            return mk_or(mk_not(lhs), rhs)
        if tok == "<->":
            self.take()
            rhs = self.implication()
            return mk_and(mk_or(mk_not(lhs), rhs), mk_or(mk_not(rhs), lhs))
        return lhs

    def disjunction(self):
        f = self.conjunction()
        while self.peek() in ("|", "||"):
            self.take()
            f = mk_or(f, self.conjunction())
        return f

    def conjunction(self):
        f = self.until()
        while self.peek() in ("&", "&&"):
            self.take()
            f = mk_and(f, self.until())
        return f

    def until(self):
        lhs = self.unary()
        tok = self.peek()
        if tok in _BINARY_TEMPORAL:
            self.take()
            rhs = self.until()  # right associative
            if tok == "U":
                return ("U", lhs, rhs)
            if tok == "R":
                return ("R", lhs, rhs)
            if tok == "W":  # f W g = g R (g | f)
                return ("R", rhs, mk_or(rhs, lhs))
            return ("U", rhs, mk_and(lhs, rhs))  # f M g = g U (f & g)
        return lhs

    def unary(self):
        tok = self.take()
        if tok is None:
            raise ValueError("unexpected end of formula")
        if tok == "!":
            return mk_not(self.unary())
        if tok == "X":
            return ("X", self.unary())
        if tok == "G":
            return ("R", FALSE, self.unary())
        if tok == "F":
            return ("U", TRUE, self.unary())
        if tok == "(":
            f = self.implication()
            if self.take() != ")":
                raise ValueError("missing ')'")
            return f
        if tok in ("true", "1"):
            return TRUE
        if tok in ("false", "0"):
            return FALSE
        if tok in _KEYWORDS or not re.match(r"[A-Za-z_]", tok):
            raise ValueError("unexpected token %r" % tok)
        return ("ap", tok)


def parse(text):
    return _Parser(_tokenize(text)).parse()


def aps_of(f, out=None):
    """Atomic propositions in order of first appearance."""
    if out is None:
        out = []
    if f[0] == "ap":
        if f[1] not in out:
            out.append(f[1])
    else:
        for g in f[1:]:
            if isinstance(g, tuple):
                aps_of(g, out)
    return out


def progress(f, letter):
    """Residual obligation after reading `letter` (dict ap -> bool)."""
    k = f[0]
    if k in ("true", "false"):
        return f
    if k == "ap":
        return TRUE if letter[f[1]] else FALSE
    if k == "not":
        return mk_not(progress(f[1], letter))
    if k == "and":
        return mk_and(progress(f[1], letter), progress(f[2], letter))
    if k == "or":
        return mk_or(progress(f[1], letter), progress(f[2], letter))
    if k == "X":
        # ("N", g): "g holds at the next position, which must exist" (strong next); the
        # wrapper keeps !X g true at the end of a trace while X g is false there
        return ("N", f[1])
    if k == "N":
        return progress(f[1], letter)
    if k == "U":  # f U g = g | (f & X(f U g))
        return mk_or(progress(f[2], letter), mk_and(progress(f[1], letter), f))
    if k == "R":  # f R g = g & (f | X(f R g))
        return mk_and(progress(f[2], letter), mk_or(progress(f[1], letter), f))
    raise ValueError(k)


def accepts_empty(f):
    """Does the empty suffix discharge the residual (finite-trace semantics)?"""
    k = f[0]
    if k == "true":
        return True
    if k in ("false", "ap", "X", "N", "U"):
        return False
    if k == "R":
        return True
    if k == "not":
        return not accepts_empty(f[1])
    if k == "and":
        return accepts_empty(f[1]) and accepts_empty(f[2])
    if k == "or":
        return accepts_empty(f[1]) or accepts_empty(f[2])
    raise ValueError(k)


def _atoms(f, out):
    if f[0] in ("ap", "X", "U", "R"):
        if f not in out:
            out.append(f)
    for g in f[1:]:
        if isinstance(g, tuple):
            _atoms(g, out)
    return out


class _Canon:
    """Truth tables over the temporal atoms of the original formula."""

    def __init__(self, formula):
        self.atoms = _atoms(formula, [])
        k = len(self.atoms)
        if k > 16:
            raise ValueError("formula has %d temporal atoms; at most 16 supported" % k)
        self.full = (1 << (1 << k)) - 1
        self.var = {}
        for i, a in enumerate(self.atoms):
            m = 0
            for row in range(1 << k):
                if (row >> i) & 1:
                    m |= 1 << row
            self.var[a] = m

    def table(self, f):
        k = f[0]
        if k == "true":
            return self.full
        if k == "false":
            return 0
        if k in ("ap", "X", "U", "R"):
            return self.var[f]
        if k == "N":  # transparent: same obligations on every non-empty continuation
            return self.table(f[1])
        if k == "not":
            return self.full ^ self.table(f[1])
        if k == "and":
            return self.table(f[1]) & self.table(f[2])
        if k == "or":
            return self.table(f[1]) | self.table(f[2])
        raise ValueError(k)


class RuleTable:
    """Deterministic monitor: trans[q][letter] -> next state or VIOLATION."""

    def __init__(self, formula, aps, trans, accepting, init=0):
        self.formula = formula
        self.aps = list(aps)  # AP names, AP k is letter bit (n_aps-1-k)
        self.trans = [list(r) for r in trans]
        self.accepting = list(accepting)
        self.init = init

    @property
    def n_states(self):
        return len(self.trans)

    @property
    def n_aps(self):
        return len(self.aps)

    def is_agent_specific(self):
        return any("#" in a for a in self.aps)

    def letter_of(self, valuation):
        n = self.n_aps
        return sum((1 if valuation[a] else 0) << (n - 1 - k) for k, a in enumerate(self.aps))

    def run(self, trace, reset_on_violation=True):
        """Monitor a finite trace (list of dict ap->bool): (#violations, final penalty flag)."""
        q, viol = self.init, 0
        for val in trace:
            nxt = self.trans[q][self.letter_of(val)]
            if nxt == VIOLATION:
                viol += 1
                q = self.init if reset_on_violation else q
            else:
                q = nxt
        return viol, (not self.accepting[q]), q

    def isomorphic(self, other):
        """Same automaton up to state renumbering (APs must be in the same order)."""
        if self.aps != other.aps or self.n_states != other.n_states:
            return False
        m = {self.init: other.init}
        todo = deque([self.init])
        while todo:
            q = todo.popleft()
            if self.accepting[q] != other.accepting[m[q]]:
                return False
            for l in range(1 << self.n_aps):
                a, b = self.trans[q][l], other.trans[m[q]][l]
                if (a == VIOLATION) != (b == VIOLATION):
                    return False
                if a == VIOLATION:
                    continue
                if a in m:
                    if m[a] != b:
                        return False
                else:
                    m[a] = b
                    todo.append(a)
        return len(set(m.values())) == len(m)


def compile_rule(formula_text, aps=None):
    """Compile an LTL formula into a RuleTable (what MakeRule does with Spot)."""
    f0 = parse(formula_text)
    names = []  # APs in textual order of first appearance (desugaring may reorder the AST)
    for tok in _tokenize(formula_text):
        if re.match(r"[A-Za-z_]", tok) and tok not in _KEYWORDS and tok not in names:
            names.append(tok)
    if aps is None:
        aps = names
    else:
        missing = [a for a in names if a not in aps]
        if missing:
            raise ValueError("formula uses APs %r not listed in aps" % missing)
    if len(aps) > MAX_APS:
        raise ValueError("at most %d atomic propositions per rule (got %d)" % (MAX_APS, len(aps)))
    for a in aps:
        if "#" in a and not a.endswith("#0"):
            raise ValueError("only the #0 agent placeholder is supported (got %r)" % a)
    canon = _Canon(f0)
    n = len(aps)
    letters = []
    for l in range(1 << n):
        letters.append({a: bool((l >> (n - 1 - k)) & 1) for k, a in enumerate(aps)})
    if canon.table(f0) == 0:
        raise ValueError("formula is unsatisfiable")
    # explore residuals
    # a residual's language = (accepts the empty suffix?, obligations on a non-empty one)
    key = lambda g: (canon.table(g), accepts_empty(g))
    index = {key(f0): 0}
    reps = [f0]
    trans = []
    todo = deque([0])
    while todo:
        q = todo.popleft()
        row = []
        for val in letters:
            g = progress(reps[q], val)
            t = key(g)
            if t[0] == 0 and not t[1]:
                row.append(VIOLATION)
                continue
            if t not in index:
                index[t] = len(reps)
                reps.append(g)
                todo.append(index[t])
            row.append(index[t])
        while len(trans) <= q:
            trans.append(None)
        trans[q] = row
    accepting = [1 if accepts_empty(r) else 0 for r in reps]
    trans, accepting = _minimise(trans, accepting, n)
    trans, accepting = _bfs_renumber(trans, accepting, n)
    if len(trans) > MAX_STATES:
        raise ValueError("monitor has %d states; at most %d supported" % (len(trans), MAX_STATES))
    return RuleTable(formula_text, aps, trans, accepting, 0)


def _minimise(trans, accepting, n_aps):
    n = len(trans)
    block = [accepting[q] for q in range(n)]
    while True:
        sig = {}
        new_block = []
        for q in range(n):
            key = (block[q],) + tuple(
                VIOLATION if t == VIOLATION else block[t] for t in trans[q])
            if key not in sig:
                sig[key] = len(sig)
            new_block.append(sig[key])
        if len(sig) == len(set(block)):
            block = new_block
            break
        block = new_block
    nb = len(set(block))
    rep = {}
    for q in range(n):
        rep.setdefault(block[q], q)
    # block of the initial state first keeps init = 0 after renumbering below
    new_trans = [None] * nb
    new_acc = [0] * nb
    for b, q in rep.items():
        new_trans[b] = [VIOLATION if t == VIOLATION else block[t] for t in trans[q]]
        new_acc[b] = accepting[q]
    init = block[0]
    if init != 0:  # swap so that init is 0
        perm = list(range(nb))
        perm[0], perm[init] = perm[init], perm[0]
        inv = {old: new for new, old in enumerate(perm)}
        new_trans = [[VIOLATION if t == VIOLATION else inv[t] for t in new_trans[perm[i]]]
                     for i in range(nb)]
        new_acc = [new_acc[perm[i]] for i in range(nb)]
    return new_trans, new_acc


def _bfs_renumber(trans, accepting, n_aps):
    order, seen = [0], {0}
    todo = deque([0])
    while todo:
        q = todo.popleft()
        for t in trans[q]:
            if t != VIOLATION and t not in seen:
                seen.add(t)
                order.append(t)
                todo.append(t)
    inv = {old: new for new, old in enumerate(order)}
    new_trans = [[VIOLATION if t == VIOLATION else inv[t] for t in trans[old]] for old in order]
    new_acc = [accepting[old] for old in order]
    return new_trans, new_acc


def satisfies(formula, trace, pos=0):
    """Direct finite-trace semantics (independent of progression; used by the tests)."""
    f = parse(formula) if isinstance(formula, str) else formula
    n = len(trace)
    k = f[0]
    if pos >= n:
        return accepts_empty(f)
    if k == "true":
        return True
    if k == "false":
        return False
    if k == "ap":
        return bool(trace[pos][f[1]])
    if k == "not":
        return not satisfies(f[1], trace, pos)
    if k == "and":
        return satisfies(f[1], trace, pos) and satisfies(f[2], trace, pos)
    if k == "or":
        return satisfies(f[1], trace, pos) or satisfies(f[2], trace, pos)
    if k == "X":
        return pos + 1 < n and satisfies(f[1], trace, pos + 1)
    if k == "U":
        for j in range(pos, n):
            if satisfies(f[2], trace, j):
                return True
            if not satisfies(f[1], trace, j):
                return False
        return False
    if k == "R":
        for j in range(pos, n):
            if not satisfies(f[2], trace, j):
                return False
            if satisfies(f[1], trace, j):
                return True
        return True
    raise ValueError(k)


class RuleMonitor:
    """Host mirror of ltl::RuleMonitor: a compiled table + weight + priority."""

    def __init__(self, table, weight, priority, on_violation="reset"):
        self.table = table
        self.weight = float(weight)
        self.priority = int(priority)
        assert on_violation in ("reset", "stay")
        self.on_violation = on_violation

    @staticmethod
    def MakeRule(formula, weight, priority, on_violation="reset"):
        return RuleMonitor(compile_rule(formula), weight, priority, on_violation)

    def IsAgentSpecific(self):
        return self.table.is_agent_specific()

    @property
    def disabled_state(self):
        """State index of "this rule does not apply to this agent" in the packed table: one extra
        absorbing, accepting state behind the automaton's own.  An agent that was not given the rule
        (BehaviorRulesMcts::AddAgentRules, src/behavior_rules_mcts.hpp:278-293) carries it in the
        rule's slots: no letter moves it, it never violates and its final penalty is 0 -- the
        kernels skip absorbing states, so such slots cost nothing."""
        return self.table.n_states

    def packed_table(self):
        """(n_states, rows, accepting) as uploaded: the automaton plus the `disabled_state`."""
        t = self.table
        rows = [list(row) for row in t.trans] + [[t.n_states] * (1 << t.n_aps)]
        return t.n_states + 1, rows, list(t.accepting) + [1]

    def to_abi(self, label_ids):
        """Pack into the C-ABI brm_rule; `label_ids` maps AP base names to label kinds."""
        from . import _abi
        t = self.table
        r = _abi.Rule()
        r.n_aps = t.n_aps
        n_states, rows, accepting = self.packed_table()
        if n_states > _abi.MAX_DFA_STATES:
            raise ValueError("rule automaton has %d states (limit %d)" % (t.n_states, _abi.MAX_DFA_STATES - 1))
        r.n_states = n_states
        r.init_state = t.init
        r.priority = self.priority
        r.on_violation = 0 if self.on_violation == "reset" else 1
        r.weight = self.weight
        for k, name in enumerate(t.aps):
            base = name.split("#")[0]
            if base not in label_ids:
                raise KeyError("unknown label %r (known: %s)" % (base, sorted(label_ids)))
            r.ap_label[k] = label_ids[base]
            r.ap_agent_specific[k] = 1 if "#" in name else 0
        stride = 1 << t.n_aps
        for q, row in enumerate(rows):
            for l, nxt in enumerate(row):
                r.trans[q * stride + l] = nxt
        for q, a in enumerate(accepting):
            r.accepting[q] = a
        return r
