"""Root-parallel partitioning and statistics merge (one process per GPU).

The reference runs one tree in one thread with a process-global std::mt19937
(/root/reference/src/behavior_rules_mcts.hpp:154).  Here every rank owns an independent
tree / rollout stream (Philox sub-stream = disjoint rollout-id range) and the only exchange
is one small sum-allreduce of the root statistics: per MCTS agent and root action the
visit count and the V-vector of summed returns -- what ReturnBestAction() and
GetEgoIntNode().GetExpectedRewards() read in the reference
(/root/reference/gridworld/grid_world_env.h:33,38).  The packed buffer is
[A][n_actions][1+V] float64, a few hundred bytes: the collective is latency-bound, so it is
a single in-place all_reduce (NCCL on GPU tensors, gloo in the CPU tests).
"""
import numpy as np


def rank_and_world(group=None):
    """(rank, world size) of this process in the root-parallel group; (0, 1) without one."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(group), dist.get_world_size(group)
    except ImportError:
        pass
    return 0, 1


# rollout ids of a host tree: every rank draws from its own 2^40-wide range
TREE_RANK_STRIDE = 1 << 40


def rollout_base(rank, decision, rollouts_per_decision, world_size=None):
    """First global rollout id of `rank` for decision number `decision`.  Ranks and decisions
    get disjoint id ranges, so no two rollouts anywhere share a Philox counter."""
    stride = 1_000_003  # > any realistic number of decisions per run; keeps ranks apart
    return (int(rank) * stride + int(decision)) * int(rollouts_per_decision)


def init_comm(engine, group=None):
    """Give `engine` (an _abi.Engine or anything with `.engine` being one) the box-wide
    communicator of the C ABI: rank 0 makes the id (brm_comm_unique_id), torch.distributed hands it
    round (plumbing), every rank calls brm_comm_init.  From then on brm_*_rollout_root_stats
    returns statistics summed over all ranks (NCCL on the engine's stream).  No-op for one rank."""
    from . import _abi
    rank, world = rank_and_world(group)
    eng = getattr(engine, "engine", engine)
    if world == 1:
        return eng
    import torch.distributed as dist
    ids = [_abi.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0, group=group)
    eng.comm_init(ids[0], rank, world)
    return eng


def shard_leaves(n_leaves, rank, world_size):
    """Contiguous share of `n_leaves` independent units (scenarios / leaves) of `rank`."""
    lo = n_leaves * rank // world_size
    hi = n_leaves * (rank + 1) // world_size
    return lo, hi


def merge_root_stats(stats, group=None):
    """In-place sum over ranks of the packed root statistics [A][n_actions][1+V].
    `stats` is a torch tensor (CUDA -> NCCL, CPU -> gloo) or a numpy array (wrapped)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return stats
    if isinstance(stats, np.ndarray):
        t = torch.from_numpy(stats)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        return stats
    dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    return stats


def expected_rewards(stats):
    """[A][n_actions][V] mean return vector per root action (NaN where never visited)."""
    s = stats.detach().cpu().numpy() if hasattr(stats, "detach") else np.asarray(stats)
    visits = s[..., 0:1]
    with np.errstate(invalid="ignore", divide="ignore"):
        return np.where(visits > 0, s[..., 1:] / visits, np.nan)


def best_actions(stats, thresholds=None):
    """Per-agent greedy root action under a thresholded lexicographic order of the mean
    return vectors (the ordering family of mvmcts::ThresUCTStatistic, un-vendored; the
    threshold semantics here: at objective k, actions whose mean is >= threshold[k] are
    considered equally good and the comparison moves on to objective k+1)."""
    q = expected_rewards(stats)
    A, n_act, V = q.shape
    thr = np.full(V, np.inf) if thresholds is None else np.asarray(thresholds, np.float64)
    out = np.zeros(A, np.int64)
    for a in range(A):
        cand = [k for k in range(n_act) if not np.isnan(q[a, k, 0])]
        if not cand:
            continue
        for v in range(V):
            vals = np.minimum(q[a, cand, v], thr[v])
            best = vals.max()
            cand = [k for k, x in zip(cand, vals) if x >= best - 1e-12]
            if len(cand) == 1:
                break
        out[a] = cand[0]
    return out
