"""Generate tests/golden/gw_golden.npz from the REFERENCE'S OWN GridWorld code
(oracle/_ref/libgw_ref.so, built from /root/reference by `make -C oracle ref`).

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_gw_golden.py
The vectors pin GridWorldState::Execute / GetTerminalReward and the label evaluators on
replayed action sequences: 96 traces x 14 steps over three scenarios (the BaseTestEnv
3-agent merge with reset- and stay-on-violation monitors, and a 4-agent variant)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import gw as ogw  # noqa: E402
from tests.gw_common import random_states  # noqa: E402

T = 14


def scenario(name):
    if name == "base3_reset":
        return ogw.base_test_env_config(3, terminal_depth=12), 3
    if name == "base3_stay":
        return ogw.base_test_env_config(3, terminal_depth=12, reset_on_violation=False), 3
    if name == "merge4_reset":
        return ogw.base_test_env_config(4, terminal_depth=10), 4
    raise KeyError(name)


def main():
    ref = ogw.GwOracle("ref")
    out = {}
    for si, name in enumerate(["base3_reset", "base3_stay", "merge4_reset"]):
        cfg, A = scenario(name)
        rng = np.random.default_rng(100 + si)
        B = 32
        agents, term = random_states(rng, B, A, p_terminal=0.1)
        # the first trace of every scenario starts from the BaseTestEnv initial state
        if A == 3:
            agents[0], term[0] = ogw.base_test_env_state()
        actions = rng.integers(0, 3, size=(B, T, A)).astype(np.uint8)
        keys = ["agents", "term", "q", "viol", "rewards", "labels"]
        acc = {k: [] for k in keys + ["term_reward", "steps"]}
        for b in range(B):
            o = ref.replay(cfg, agents[b], term[b], 0, actions[b])
            for k in keys + ["term_reward"]:
                acc[k].append(o[k])
            acc["steps"].append(o["steps"])
        out[name + "/agents0"] = agents
        out[name + "/term0"] = term
        out[name + "/actions"] = actions
        for k, v in acc.items():
            out[name + "/" + k] = np.array(v)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "gw_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
