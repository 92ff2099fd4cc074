#!/usr/bin/env python
"""Randomised parity campaign on the CPU: the host build of the device work items
(tests/cpp/cw_emul.cpp over csrc/cw_core.cuh -- bit for bit what the GPU computes,
tests/test_cw_gpu.py::test_gpu_bits_equal_the_host_compiled_work_items) against the FP64 oracle,
rollout by rollout, over freshly seeded merge / highway / intersection scenarios.  Every discrete
output must be EQUAL (tests/cw_common.py::compare_rollouts); prints the smallest threshold margin
the oracle saw, i.e. how far the closest comparison was from being decided by rounding.
    python tools/parity_campaign.py [first_seed] [seconds]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cw as ocw  # noqa: E402
from planner_rules_mcts_b200 import scenarios  # noqa: E402
from tests.cw_common import compare_rollouts, oracle_config  # noqa: E402
from tests.cw_emul import CwEmul  # noqa: E402

orc, em = ocw.CwOracle(), CwEmul()
FL = lambda n: np.full(n, ocw.PRESENT | ocw.VALID, np.uint8)  # noqa: E731
makers = [("merge", lambda s: scenarios.merge_scenario(seed=s)),
          ("highway", lambda s: scenarios.highway_scenario(seed=s)),
          ("intersection", lambda s: scenarios.intersection_scenario(seed=1000, scenario_id=s))]
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 0
budget = float(sys.argv[2]) if len(sys.argv) > 2 else 600.0
t0, total, mismatches, min_margin = time.time(), 0, 0, 1e9
while time.time() - t0 < budget:
    for name, mk in makers:
        scn = mk(seed)
        cfg = oracle_config(scn)
        A = len(scn.agents)
        n = 120 if name == "intersection" else 400
        args = (scn.initial_state, FL(A), scn.params.HORIZON, 77 + seed, 1000, n, 30, scn.params.DISCOUNT_FACTOR)
        ref = orc.rollout_detail(cfg, *args)
        got = em.rollout_detail(scn, *args)
        try:
            compare_rollouts(got, ref, "%s seed %d" % (name, seed))
        except AssertionError as e:
            mismatches += 1
            print("MISMATCH", name, seed, str(e)[:300], flush=True)
        min_margin = min(min_margin, float(ref["margin"].min()))
        total += n
    seed += 1
print("rollouts %d, scenarios %d, mismatches %d, smallest oracle margin %.3g, %.0f s"
      % (total, 3 * (seed - (int(sys.argv[1]) if len(sys.argv) > 1 else 0)), mismatches, min_margin, time.time() - t0))
sys.exit(1 if mismatches else 0)
