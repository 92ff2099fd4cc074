"""Host-side multi-agent MCTS that drives the device engine.

The tree stays on the host, as north_star asks; every iteration expands one joint action with
one batched `Execute` and evaluates the new node with ONE device launch of
`rollouts_per_expansion` random rollouts (the mvmcts::RandomHeuristic call of the reference's
iteration, src/behavior_rules_mcts.hpp:188-189,232).  Selection is decoupled UCT: every agent
keeps its own per-action visit counts and value-vector sums and picks its action
independently (untried actions first, then UCB on normalised values); vector values are
compared in thresholded lexicographic order.  The reference's statistics classes
(`ThresUCTStatistic`, `ThresGreedyStatistic`; un-vendored lexmamcts@24ab42ca) are not available,
so the exact tie-breaking and exploration schedule are this module's own ("parity unpinned");
what IS pinned are the planner-level behaviours of test/single_agent_test.cpp:229-270.
"""
import math

import numpy as np

from . import parallel


class _Node:
    __slots__ = ("state", "terminal", "n_actions", "children", "N", "Q", "value", "term_value")

    def __init__(self, state, terminal, n_actions, V):
        self.state = state
        self.terminal = bool(terminal)
        self.n_actions = [int(k) for k in n_actions]
        self.children = {}
        self.N = [np.zeros(k) for k in self.n_actions]
        self.Q = [np.zeros((k, V)) for k in self.n_actions]
        self.value = np.zeros((len(self.n_actions), V))  # the heuristic's estimate at expansion
        self.term_value = None                             # terminal nodes: GetTerminalReward, cached


class MctsParameters:
    """The mvmcts::MvmctsParameters fields the reference sets (src/util.cpp:23-89)."""

    def __init__(self, V, discount_factor=0.9, exploration_constant=1.42, lower_bound=None, upper_bound=None,
                 threshold=None, max_rollout_steps=30, rollouts_per_expansion=16, seed=1000, coop_factor=0.0,
                 statistic="uct", decay1=0.8, decay2=0.12, batch_size=1, rank=None):
        self.rank = rank                          # root-parallel rank (None: from torch.distributed, else 0)
        self.DISCOUNT_FACTOR = discount_factor
        self.COOP_FACTOR = coop_factor            # src/util.cpp:28-29
        # "uct": ThresUCTStatistic; "greedy": ThresGreedyStatistic (epsilon-greedy on the thresholded
        # order, epsilon = DECAY1 / (1 + DECAY2 * visits of the node); gridworld/base_test_env.cpp:69-70)
        assert statistic in ("uct", "greedy")
        self.statistic, self.DECAY1, self.DECAY2 = statistic, decay1, decay2
        # iterations that descend the tree before the device is called: 1 = the reference's loop;
        # > 1: leaf-parallel rounds with a virtual loss (one batched Execute + one rollout launch
        # for all of them, INTEGRATION.md level 2)
        self.batch_size = max(1, int(batch_size))
        self.EXPLORATION_CONSTANT = exploration_constant
        lb = [-1.0] * V
        lb[-1] = -1000.0                      # src/util.cpp:58-59
        self.LOWER_BOUND = np.asarray(lower_bound if lower_bound is not None else lb, np.float64)
        self.UPPER_BOUND = np.asarray(upper_bound if upper_bound is not None else [0.0] * V, np.float64)
        self.THRESHOLD = None if threshold is None else np.asarray(threshold, np.float64)
        self.MAX_NUMBER_OF_ITERATIONS = max_rollout_steps
        self.rollouts_per_expansion = rollouts_per_expansion
        self.seed = seed


class Mcts:
    """`Search(state, max_iterations)` / `ReturnBestAction()` / `GetRootStats()` -- the call
    surface BehaviorRulesMcts::Plan uses on mvmcts::Mvmcts (src/behavior_rules_mcts.hpp:232-249)."""

    def __init__(self, engine, params: MctsParameters):
        self.engine = engine
        self.p = params
        self.root = None
        self.iterations = 0
        # root-parallel search: every rank grows its own tree from its own random streams (Philox
        # rollout ids and the host generator of the epsilon-greedy statistic); identical streams
        # would make the merged statistics N copies of one tree
        self.rank = parallel.rank_and_world()[0] if params.rank is None else int(params.rank)
        self._next_rollout = self.rank * parallel.TREE_RANK_STRIDE
        # selection compares normalised values (+ exploration bonus), so the thresholds are
        # normalised with the same bounds
        span = np.maximum(self.p.UPPER_BOUND - self.p.LOWER_BOUND, 1e-9)
        self._thr_norm = None if self.p.THRESHOLD is None else (self.p.THRESHOLD - self.p.LOWER_BOUND) / span
        self._rng = np.random.default_rng([int(self.p.seed), self.rank])

    # -- helpers -----------------------------------------------------------------------
    def _make_node(self, state):
        return _Node(state, self.engine.is_terminal(state)[0], self.engine.get_num_actions(state)[:, 0],
                     self.engine.V)

    def _lex_better(self, a, b):
        """a better than b in thresholded lexicographic order."""
        thr = self._thr_norm
        for k in range(len(a)):
            x, y = a[k], b[k]
            if thr is not None:
                x, y = min(x, thr[k]), min(y, thr[k])
            if abs(x - y) > 1e-12:
                return x > y
        return False

    def _select(self, node):
        ja = []
        span = np.maximum(self.p.UPPER_BOUND - self.p.LOWER_BOUND, 1e-9)
        for a, k in enumerate(node.n_actions):
            untried = np.flatnonzero(node.N[a] == 0)
            if len(untried):
                ja.append(int(untried[0]))
                continue
            total = node.N[a].sum()
            mean = node.Q[a] / node.N[a][:, None]
            norm = (mean - self.p.LOWER_BOUND) / span
            if self.p.statistic == "greedy":
                if self._rng.random() < self.p.DECAY1 / (1.0 + self.p.DECAY2 * total):
                    ja.append(int(self._rng.integers(k)))
                    continue
                best = 0
                for act in range(1, k):
                    if self._lex_better(norm[act], norm[best]):
                        best = act
                ja.append(best)
                continue
            bonus = np.sqrt(2.0 * math.log(total) / node.N[a])
            ucb = norm + bonus[:, None] * np.asarray(self.p.EXPLORATION_CONSTANT, np.float64)
            best = 0
            for act in range(1, k):
                if self._lex_better(ucb[act], ucb[best]):
                    best = act
            ja.append(best)
        return tuple(ja)

    def _evaluate(self, node):
        """RandomHeuristic: mean discounted return of random rollouts from the node."""
        n = self.p.rollouts_per_expansion
        out = self.engine.rollout(node.state, n, seed=self.p.seed, rollout_base=self._next_rollout,
                                  max_steps=self.p.MAX_NUMBER_OF_ITERATIONS,
                                  discount_factor=self.p.DISCOUNT_FACTOR, first_discount_exp=0,
                                  coop_factor=self.p.COOP_FACTOR)
        self._next_rollout += n
        return out["returns_sum"][0] / n  # [M][V]

    def _coop(self, r):
        """Cooperative reward [M][V]: own + COOP_FACTOR * the others' (applied to step and terminal
        rewards here, to rollout returns on the device)."""
        c = self.p.COOP_FACTOR
        return r if c == 0.0 else r + c * (r.sum(axis=0, keepdims=True) - r)

    # -- search ----------------------------------------------------------------------------
    def Search(self, state, max_iterations):
        self.root = self._make_node(state)
        if self.root.terminal:
            raise RuntimeError("Current state is terminal! Aborting!")  # src/behavior_rules_mcts.hpp:208-209
        if self.p.batch_size > 1:
            return self._search_batched(max_iterations)
        gamma = self.p.DISCOUNT_FACTOR
        M, V = self.engine.M, self.engine.V
        for _ in range(max_iterations):
            node, path = self.root, []
            while True:
                if node.terminal:
                    value = self._coop(self.engine.terminal_reward(node.state)[:, :, 0].astype(np.float64))
                    break
                ja = self._select(node)
                if ja in node.children:
                    child, r = node.children[ja]
                    path.append((node, ja, r))
                    node = child
                    continue
                nxt, rew = self.engine.execute(node.state, np.array(ja, np.uint8).reshape(M, 1))
                child = self._make_node(nxt)
                r = self._coop(rew[:, :, 0].astype(np.float64))
                node.children[ja] = (child, r)
                path.append((node, ja, r))
                value = self._evaluate(child)
                child.value = value
                break
            for nd, ja, r in reversed(path):
                value = r + gamma * value
                for a in range(M):
                    nd.N[a][ja[a]] += 1
                    nd.Q[a][ja[a]] += value[a]
            self.iterations += 1
        return self

    def _search_batched(self, max_iterations):
        """Leaf-parallel rounds: `batch_size` descents per round, kept apart by a virtual loss (each
        descent counts its visits at once, with the lower bound as value, until its real return is
        known); the new nodes of a round are created by ONE batched Execute and evaluated by ONE
        rollout launch.  Descents that ask for the same expansion share it."""
        eng, p = self.engine, self.p
        gamma, M = p.DISCOUNT_FACTOR, eng.M
        vloss = p.LOWER_BOUND.astype(np.float64)
        batch_type = type(self.root.state)
        done = 0
        while done < max_iterations:
            k = min(p.batch_size, max_iterations - done)
            descents = []          # (path, ("term", node) | ("exp", index into expansions))
            expansions, index = [], {}
            for _ in range(k):
                node, path = self.root, []
                while True:
                    if node.terminal:
                        descents.append((path, ("term", node)))
                        break
                    ja = self._select(node)
                    for a in range(M):                      # virtual loss
                        node.N[a][ja[a]] += 1
                        node.Q[a][ja[a]] += vloss
                    if ja in node.children:
                        child, r = node.children[ja]
                        path.append((node, ja, r))
                        node = child
                        continue
                    key = (id(node), ja)
                    if key not in index:
                        index[key] = len(expansions)
                        expansions.append((node, ja))
                    path.append((node, ja, None))
                    descents.append((path, ("exp", index[key])))
                    break
            new_nodes = []
            if expansions:
                parents = batch_type.concat([nd.state for nd, _ in expansions])
                acts = np.array([ja for _, ja in expansions], np.uint8).T.copy()       # [M][n]
                nxt, rew = eng.execute(parents, acts)
                term = eng.is_terminal(nxt)
                nact = eng.get_num_actions(nxt)
                tval = eng.terminal_reward(nxt).astype(np.float64)
                n = p.rollouts_per_expansion
                out = eng.rollout(nxt, n, seed=p.seed, rollout_base=self._next_rollout,
                                  max_steps=p.MAX_NUMBER_OF_ITERATIONS, discount_factor=p.DISCOUNT_FACTOR,
                                  first_discount_exp=0, coop_factor=p.COOP_FACTOR)
                self._next_rollout += n * len(expansions)
                for i, (nd, ja) in enumerate(expansions):
                    child = _Node(nxt.select(np.array([i])), term[i], nact[:, i], eng.V)
                    child.value = out["returns_sum"][i] / n
                    child.term_value = self._coop(tval[:, :, i])
                    r = self._coop(rew[:, :, i].astype(np.float64))
                    nd.children[ja] = (child, r)
                    new_nodes.append((child, r))
            for path, (kind, what) in descents:
                if kind == "term":
                    if what.term_value is None:
                        what.term_value = self._coop(eng.terminal_reward(what.state)[:, :, 0].astype(np.float64))
                    value = what.term_value
                else:
                    child, r_new = new_nodes[what]
                    value = child.value
                for nd, ja, r in reversed(path):
                    r = r_new if r is None else r
                    value = r + gamma * value
                    for a in range(M):                      # replace the virtual loss by the return
                        nd.Q[a][ja[a]] += value[a] - vloss
                self.iterations += 1
            done += k
        return self

    def GetRootStats(self):
        """Packed root statistics [M][n_actions][1+V] (what ranks allreduce)."""
        M, V = self.engine.M, self.engine.V
        k = max(self.root.n_actions)
        s = np.zeros((M, k, 1 + V))
        for a in range(M):
            s[a, :self.root.n_actions[a], 0] = self.root.N[a]
            s[a, :self.root.n_actions[a], 1:] = self.root.Q[a]
        return s

    def ReturnBestAction(self, merge=True):
        stats = self.GetRootStats()
        if merge:
            parallel.merge_root_stats(stats)  # root-parallel: sum over ranks
        return [int(x) for x in parallel.best_actions(stats, self.p.THRESHOLD)]

    def NumIterations(self):
        return self.iterations

    def GetTree(self, value_idx=0):
        """BehaviorRulesMcts::GetTree / DfsTree (src/behavior_rules_mcts.hpp:431-460): one
        (value, polyline) pair per leaf of the search tree -- the ego's positions along the path and
        the ego's rewards summed along it plus the leaf's value estimate, component `value_idx`."""
        lines = []

        def dfs(node, value, pts):
            xy = node.state.agents[0, 0, :2] if hasattr(node.state, "flags") else node.state.agents[0, 0, :1]
            pts = pts + [tuple(float(c) for c in xy)]
            if node.children:
                for child, r in node.children.values():
                    value = value + float(r[0][value_idx])   # the reference accumulates across siblings too
                    dfs(child, value, pts)
            else:
                lines.append((value + float(node.value[0][value_idx]), pts))

        if self.root is not None:
            dfs(self.root, 0.0, [])
        return lines
