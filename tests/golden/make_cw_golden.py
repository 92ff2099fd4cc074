"""Generate tests/golden/cw_golden.npz with the FP64 continuous-world oracle
(oracle/cw_oracle.cpp).  NOTE: this oracle restates un-vendored dependencies ("parity
unpinned", see oracle/cw_oracle.h); these vectors freeze ITS behaviour so that neither the
oracle nor the CUDA kernels can drift silently -- they are not outputs of the reference.
    python tests/golden/make_cw_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import cw as ocw  # noqa: E402
from planner_rules_mcts_b200 import scenarios  # noqa: E402
from tests.cw_common import oracle_config  # noqa: E402

T, B = 20, 16
SCENARIOS = {
    "highway": lambda: scenarios.highway_scenario(seed=1000),
    "merge": lambda: scenarios.merge_scenario(seed=1000),
    "intersection": lambda: scenarios.intersection_scenario(seed=1000, scenario_id=3),
}


def main():
    orc = ocw.CwOracle()
    out = {}
    for si, (name, make) in enumerate(SCENARIOS.items()):
        scn = make()
        cfg = oracle_config(scn)
        A, M = len(scn.agents), scn.n_mcts_agents
        rng = np.random.default_rng(500 + si)
        n_act = len(scn.agents[0].primitives)
        acts = rng.integers(0, n_act, size=(B, T, M)).astype(np.uint8)
        fl = np.full(A, 3, np.uint8)
        keys = ["states", "flags", "terminal", "q", "viol", "rewards", "labels", "frenet", "margin"]
        acc = {k: [] for k in keys + ["term_reward", "steps"]}
        for b in range(B):
            o = orc.replay(cfg, scn.initial_state, fl, scn.params.HORIZON, acts[b])
            for k in keys + ["term_reward"]:
                acc[k].append(o[k])
            acc["steps"].append(o["steps"])
        out[name + "/actions"] = acts
        for k, v in acc.items():
            a = np.array(v)
            if a.dtype == np.float64 and k != "margin":
                a = a.astype(np.float64)
            out[name + "/" + k] = a
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cw_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
