"""N > 1 host logic on CPU: two gloo ranks, each with its own leaf results, merge their root
statistics with one allreduce; every rank must end with the same totals and the same best
actions, and the rollout-id ranges of ranks / decisions must be disjoint."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np

from planner_rules_mcts_b200 import parallel

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def host_root_stats(first, returns_sum, n, A, n_act, V):
    """numpy restatement of K5 (brm_root_stats_reduce) used as the expectation."""
    s = np.zeros((A, n_act, 1 + V))
    for l in range(first.shape[0]):
        for a in range(A):
            s[a, first[l, a], 0] += n
            s[a, first[l, a], 1:] += returns_sum[l, a]
    return s


WORKER = textwrap.dedent("""
    import os, sys
    import numpy as np
    import torch
    import torch.distributed as dist
    sys.path.insert(0, %(root)r)
    from planner_rules_mcts_b200 import parallel
    from tests.test_parallel_gloo import host_root_stats
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo")
    A, n_act, V, L, n = 4, 3, 3, 24, 128
    rng = np.random.default_rng(100 + rank)           # every rank owns different leaves
    first = rng.integers(0, n_act, size=(L, A))
    ret = rng.normal(size=(L, A, V)) * n
    local = host_root_stats(first, ret, n, A, n_act, V)
    merged = torch.from_numpy(local.copy())
    parallel.merge_root_stats(merged)
    # expectation: recompute every rank's contribution locally
    want = np.zeros_like(local)
    for r in range(world):
        g = np.random.default_rng(100 + r)
        f = g.integers(0, n_act, size=(L, A)); x = g.normal(size=(L, A, V)) * n
        want += host_root_stats(f, x, n, A, n_act, V)
    assert np.allclose(merged.numpy(), want, rtol=1e-12), "allreduce(sum) of root stats is wrong"
    assert merged.numpy()[..., 0].sum() == world * L * n * A
    best = parallel.best_actions(merged, thresholds=[-0.5, -0.5, np.inf])
    gathered = [None] * world
    dist.all_gather_object(gathered, best.tolist())
    assert all(g == gathered[0] for g in gathered), "ranks disagree on the best action"
    # numpy input path
    arr = local.copy()
    parallel.merge_root_stats(arr)
    assert np.allclose(arr, want, rtol=1e-12)
    lo, hi = parallel.shard_leaves(4096, rank, world)
    spans = [None] * world
    dist.all_gather_object(spans, (lo, hi))
    assert spans[0][0] == 0 and spans[-1][1] == 4096 and all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
    # every rank grows its own tree: Philox rollout ids and the host generator differ per rank
    # (identical streams would make the merged statistics `world` copies of one tree)
    from planner_rules_mcts_b200.mcts import Mcts, MctsParameters
    assert parallel.rank_and_world() == (rank, world)
    m = Mcts(None, MctsParameters(3, statistic="greedy"))
    streams = [None] * world
    dist.all_gather_object(streams, (m._next_rollout, float(m._rng.random())))
    assert len({s[0] for s in streams}) == world and len({s[1] for s in streams}) == world, streams
    assert all(abs(streams[i][0] - streams[j][0]) >= 2 ** 40 for i in range(world) for j in range(i))
    assert parallel.rollout_base(rank, 3, 1000) != parallel.rollout_base((rank + 1) %% world, 3, 1000)
    dist.barrier()
    dist.destroy_process_group()
    print("rank", rank, "ok")
""")


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_two_rank_gloo_merge(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % dict(root=ROOT))
    port = free_port()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=180)[0] for p in procs]
    for rank, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, "rank %d failed:\n%s" % (rank, o)
        assert "rank %d ok" % rank in o


def test_rollout_id_ranges_are_disjoint():
    n = 262144
    seen = []
    for rank in range(8):
        for decision in (0, 1, 2, 500):
            b = parallel.rollout_base(rank, decision, n)
            seen.append((b, b + n))
    seen.sort()
    assert all(seen[i][1] <= seen[i + 1][0] for i in range(len(seen) - 1))


def test_best_actions_thresholded_lexicographic():
    # agent 0: action 1 is best on objective 0 unless the threshold makes 0 and 1 tie, then
    # objective 1 decides
    stats = np.zeros((1, 3, 3))
    stats[0, :, 0] = 10
    stats[0, 0, 1:] = [-2.0, -1.0]   # mean (-0.2, -0.1)
    stats[0, 1, 1:] = [-1.0, -9.0]   # mean (-0.1, -0.9)
    stats[0, 2, 1:] = [-30.0, 0.0]   # mean (-3.0,  0.0)
    assert parallel.best_actions(stats).tolist() == [1]
    assert parallel.best_actions(stats, thresholds=[-0.25, np.inf]).tolist() == [0]
    q = parallel.expected_rewards(stats)
    assert np.allclose(q[0, 2], [-3.0, 0.0])
