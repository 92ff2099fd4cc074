"""Out-of-bounds writes: compute-sanitizer is closed on the GPU pool, so the kernels are checked
with guard bands instead -- every output array of a launch is carved out of one device buffer
filled with a canary pattern, with gaps before, between and after the arrays; after the launch
the gaps must be untouched and the arrays must equal a run with ordinary allocations."""
import numpy as np
import pytest
import torch

import planner_rules_mcts_b200 as brm
from planner_rules_mcts_b200 import scenarios

pytestmark = pytest.mark.gpu

CANARY = 0xA5
GAP = 4096  # bytes


class Arena:
    """Carves tensors out of one canary-filled buffer and checks the space between them."""

    def __init__(self, nbytes):
        self.buf = torch.full((nbytes,), CANARY, dtype=torch.uint8, device="cuda")
        self.off = GAP
        self.used = []

    def take(self, shape, dtype):
        n = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
        assert self.off + n + GAP <= self.buf.numel()
        t = self.buf[self.off:self.off + n].view(dtype).view(*shape)
        self.used.append((self.off, self.off + n))
        self.off = (self.off + n + GAP + 255) // 256 * 256
        return t

    def check(self):
        mask = torch.ones(self.buf.numel(), dtype=torch.bool, device="cuda")
        for a, b in self.used:
            mask[a:b] = False
        bad = (self.buf[mask] != CANARY).nonzero()
        assert bad.numel() == 0, "guard band overwritten at %d byte(s)" % bad.numel()


def cw_case(name):
    if name == "merge":
        return [scenarios.merge_scenario()], 37, 33
    if name == "highway":           # simulated-only agents in shared-memory slots
        return [scenarios.highway_scenario()], 5, 200
    return [scenarios.intersection_scenario(scenario_id=k) for k in range(7)], None, 29   # several scenarios


@pytest.mark.parametrize("name", ["merge", "highway", "intersections"])
def test_cw_rollout_writes_stay_inside_their_arrays(built, name):
    scns, copies, n = cw_case(name)
    eng = brm.RulesMctsEngine(scns)
    leaves = eng.initial_batch() if copies is None else eng.initial_batch([0]).select(np.zeros(copies, np.int64))
    L, N = leaves.n, leaves.n * n
    A, M, V, R = eng.A, eng.M, eng.V, eng.R
    ref = eng.rollout(leaves, n, seed=5, max_steps=15, detail=True)
    ar = Arena(64 << 20)
    out = dict(returns_sum=ar.take((L, M, V), torch.float64), viol_sum=ar.take((L, M, R), torch.int64),
               agent_steps=ar.take((L,), torch.int64), ro_returns=ar.take((N, M, V), torch.float64),
               ro_viol=ar.take((N, M, R), torch.uint8), ro_q=ar.take((N, M, R), torch.uint8),
               ro_agents=ar.take((N, A, 4), torch.float64), ro_flags=ar.take((N, A), torch.uint8),
               ro_steps=ar.take((N,), torch.int32))
    d_leaves = leaves.cuda(0)
    torch.cuda.synchronize()   # the engine runs on its own stream
    eng.rollout(d_leaves, n, seed=5, max_steps=15, out=out)
    torch.cuda.synchronize()
    ar.check()
    for k, v in out.items():
        got = v.cpu().numpy()
        assert np.array_equal(got.view(ref[k].dtype).reshape(ref[k].shape), ref[k]), k


def test_cw_execute_and_evaluate_rules_writes_stay_inside(built):
    scn = scenarios.highway_scenario()
    eng = brm.RulesMctsEngine([scn])
    b = eng.initial_batch([0]).select(np.zeros(101, np.int64))       # not a multiple of the warp's states
    acts = np.random.default_rng(1).integers(0, len(scn.agents[0].primitives), size=(eng.M, b.n)).astype(np.uint8)
    ref, ref_rew = eng.execute(b, acts)
    A, M, V, R, n = eng.A, eng.M, eng.V, eng.R, b.n
    for call in ("execute", "evaluate_rules"):
        ar = Arena(8 << 20)
        out = brm.CwBatch(ar.take((n,), torch.int32), ar.take((A, n, 4), torch.float64), ar.take((A, n), torch.uint8),
                          ar.take((n,), torch.int32), ar.take((n,), torch.uint8), ar.take((M, R, n), torch.uint8),
                          ar.take((M, R, n), torch.int32))
        rew = ar.take((M, V, n), torch.float64)
        d_in = b.cuda(0)
        s_in, s_out = d_in.to_abi(), out.to_abi()
        import ctypes as C
        from planner_rules_mcts_b200 import _abi
        d_acts = torch.from_numpy(acts).cuda()
        torch.cuda.synchronize()
        if call == "execute":
            _abi.check(eng.lib.brm_cw_execute(eng.engine.handle, C.byref(s_in), d_acts.data_ptr(), C.byref(s_out),
                                              rew.data_ptr()))
        else:
            _abi.check(eng.lib.brm_cw_evaluate_rules(eng.engine.handle, C.byref(s_in), C.byref(s_out), rew.data_ptr()))
        torch.cuda.synchronize()
        ar.check()
        if call == "execute":
            assert np.array_equal(out.agents.cpu().numpy(), ref.agents) and np.array_equal(out.q.cpu().numpy(), ref.q)
            assert np.array_equal(rew.cpu().numpy(), ref_rew)


def test_gw_rollout_and_execute_writes_stay_inside(built):
    eng = brm.GridWorldEngine()
    leaves = eng.base_test_env_batch(19)
    n, L = 77, 19
    N, A, V, R = L * n, eng.A, eng.V, eng.R
    ref = eng.rollout(leaves, n, seed=9, detail=True)
    ar = Arena(32 << 20)
    out = dict(returns_sum=ar.take((L, A, V), torch.float64), viol_sum=ar.take((L, A, R), torch.int64),
               agent_steps=ar.take((L,), torch.int64), ro_returns=ar.take((N, A, V), torch.float64),
               ro_viol=ar.take((N, A, R), torch.uint8), ro_q=ar.take((N, A, R), torch.uint8),
               ro_agents=ar.take((N, A, 4), torch.int32), ro_steps=ar.take((N,), torch.int32))
    d_leaves = leaves.cuda(0)
    torch.cuda.synchronize()
    eng.rollout(d_leaves, n, seed=9, out=out)
    torch.cuda.synchronize()
    ar.check()
    for k, v in out.items():
        assert np.array_equal(v.cpu().numpy().view(ref[k].dtype).reshape(ref[k].shape), ref[k]), k
    # Execute
    ar = Arena(4 << 20)
    m = 1001
    b = eng.base_test_env_batch(m)
    acts = np.random.default_rng(2).integers(0, 3, size=(A, m)).astype(np.uint8)
    ref_b, ref_rew = eng.execute(b, acts)
    ob = brm.GwBatch(ar.take((A, m, 4), torch.int32), ar.take((m,), torch.uint8), ar.take((m,), torch.int32),
                     ar.take((A, R, m), torch.uint8), ar.take((A, R, m), torch.int32))
    ob.viol.zero_()
    rew = ar.take((A, V, m), torch.float64)
    import ctypes as C
    from planner_rules_mcts_b200 import _abi
    d_in, d_acts = b.cuda(0), torch.from_numpy(acts).cuda()
    s_in, s_out = d_in.to_abi(), ob.to_abi()
    torch.cuda.synchronize()
    _abi.check(eng.lib.brm_gw_execute(eng.engine.handle, C.byref(s_in), d_acts.data_ptr(), C.byref(s_out),
                                      rew.data_ptr()))
    torch.cuda.synchronize()
    ar.check()
    assert np.array_equal(ob.agents.cpu().numpy(), ref_b.agents) and np.array_equal(rew.cpu().numpy(), ref_rew)


def test_the_guard_bands_do_catch_an_overrun(built):
    """Self-test of the mechanism: an output array that is one rollout too short must trip it."""
    eng = brm.GridWorldEngine()
    leaves = eng.base_test_env_batch(3)
    n = 40
    N = 3 * n
    ar = Arena(1 << 20)
    out = dict(returns_sum=ar.take((3, eng.A, eng.V), torch.float64), ro_steps=ar.take((N - 1,), torch.int32))
    d_leaves = leaves.cuda(0)
    torch.cuda.synchronize()
    eng.rollout(d_leaves, n, seed=9, out=out)
    torch.cuda.synchronize()
    with pytest.raises(AssertionError, match="guard band"):
        ar.check()
