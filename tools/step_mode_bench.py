#!/usr/bin/env python
"""Step-mode kernels (K2 gw_execute_kernel, K4 cw_pool_kernel<EXECUTE>) against the HBM
roofline: one Execute for n device-resident states, SoA in -> SoA out.  These kernels are the
StateInterface surface for host-side tree expansion (not the benchmark's critical path); they
stream every state once, so their bound is HBM.  Algorithmic bytes per state are counted from
the arrays the kernel must read and write.  Prints one JSON object per kernel."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import planner_rules_mcts_b200 as brm  # noqa: E402
from planner_rules_mcts_b200 import scenarios  # noqa: E402

PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def timeit(fn, stream, iters=20, warm=3):
    for _ in range(warm):
        fn()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    evs = []
    for i in range(iters):
        flush.fill_(i & 255)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        fn()
        b.record(stream)
        evs.append((a, b))
    torch.cuda.synchronize()
    return float(np.median([a.elapsed_time(b) for a, b in evs]))


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 22
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    out = []
    # ---- GridWorld K2
    eng = brm.GridWorldEngine(device=0, cuda_stream=stream.cuda_stream)
    A, R, V = eng.A, eng.R, eng.V
    rng = np.random.default_rng(0)
    agents = np.zeros((n, A, 4), np.int32)
    agents[:, :, 0] = rng.integers(0, 12, size=(n, A))
    agents[:, :, 1] = rng.integers(-1, 2, size=(n, A))
    agents[:, :, 3] = rng.integers(1, 4, size=(n, A))
    agents[:, :, 2] = np.where(agents[:, :, 0] >= 8, 0, agents[:, :, 3])
    batch = eng.make_batch(agents, track_violations=False).cuda(0)
    acts = torch.from_numpy(rng.integers(0, 3, size=(A, n)).astype(np.uint8)).cuda()
    ms = timeit(lambda: eng.execute(batch, acts), stream)
    rd = A * 16 + 1 + 4 + A * R + A
    wr = A * 16 + 1 + 4 + A * R + A * V * 8
    gbs = n * (rd + wr) / (ms * 1e-3) / 1e9
    out.append(dict(kernel="gw_execute_kernel<3>", states=n, ms=ms, bytes_per_state=rd + wr,
                    achieved_gbs=gbs, peak_gbs=PEAK, frac=gbs / PEAK, agent_steps_per_s=n * A / (ms * 1e-3)))
    del batch, acts
    # ---- continuous world K4 (config M scenario)
    scn = scenarios.merge_scenario(seed=1000)
    ce = brm.RulesMctsEngine([scn], device=0, cuda_stream=stream.cuda_stream)
    A, M, R, V = ce.A, ce.M, ce.R, ce.V
    n2 = n // 2
    st = np.broadcast_to(scn.initial_state, (n2, A, 4)).copy()
    st[:, :, 0] += rng.uniform(-20, 20, size=(n2, A))
    cb = brm.CwBatch.zeros(A, M, R, n2, track_violations=False)
    cb.agents[:] = st.transpose(1, 0, 2)
    cb.flags[:] = 3
    cb.horizon[:] = 20
    cb.q[:] = ce.initial_q()[:, :, None]
    dcb = cb.cuda(0)
    acts = torch.from_numpy(rng.integers(0, 3, size=(M, n2)).astype(np.uint8)).cuda()
    ms = timeit(lambda: ce.execute(dcb, acts), stream)
    # FP64 states: 32 bytes per agent in and out, rewards are doubles
    rd = A * 32 + A + 4 + 1 + 4 + M * R + M
    wr = A * 32 + A + 4 + 1 + 4 + M * R + M * V * 8
    gbs = n2 * (rd + wr) / (ms * 1e-3) / 1e9
    out.append(dict(kernel="cw_pool_kernel<EXECUTE> (config M scenario)", states=n2, ms=ms, bytes_per_state=rd + wr,
                    achieved_gbs=gbs, peak_gbs=PEAK, frac=gbs / PEAK, agent_steps_per_s=n2 * A / (ms * 1e-3)))
    for o in out:
        print(json.dumps(o))


if __name__ == "__main__":
    main()
