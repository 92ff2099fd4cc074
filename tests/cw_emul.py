"""TEST INFRASTRUCTURE -- ctypes access to tests/cpp/cw_emul.cpp: the work-item functions the
continuous-world CUDA kernels are made of (planner-rules-mcts_b200/csrc/cw_core.cuh), compiled for
the host and run item by item on the CPU.  Lets the `-m "not gpu"` suite compare the device
arithmetic with the oracle where there is no GPU.  Outputs use the layouts of the C ABI
(brm_cw_trace / brm_cw_rollout_out), so the same comparison helpers serve both."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "cw_emul.cpp")
LIB = os.path.join(ROOT, "build", "libcw_emul.so")
CSRC = os.path.join(ROOT, "planner-rules-mcts_b200", "csrc")


def build():
    deps = [SRC, os.path.join(CSRC, "cw_core.cuh"), os.path.join(CSRC, "cw_blob_build.h"),
            os.path.join(ROOT, "include", "brm.h")]
    if os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in deps):
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-mfma", "-Wall", "-Wno-unused-function",
                    "-shared", "-fPIC", "-I", os.path.join(ROOT, "include"), "-I", CSRC, "-x", "c++", SRC,
                    "-o", LIB], check=True)
    return LIB


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


class CwEmul:
    def __init__(self):
        self.lib = C.CDLL(build())
        self.lib.cwe_last_error.restype = C.c_char_p

    def _check(self, rc):
        if rc < 0:
            raise RuntimeError(self.lib.cwe_last_error().decode())
        return rc

    def sincos(self, x):
        s, c = C.c_double(0), C.c_double(0)
        self.lib.cwe_sincos(C.c_double(x), C.byref(s), C.byref(c))
        return s.value, c.value

    def frenet(self, scn, x, y):
        out = np.zeros(3)
        abi = scn.to_abi()
        self._check(self.lib.cwe_frenet(C.byref(abi), C.c_double(x), C.c_double(y), _p(out, C.c_double)))
        return int(out[0]), out[1], out[2]

    def classify(self, scn, x, y, th, agent=0):
        """grid classification of the drivable-area test at a pose: 0 / 1 = decided (and equal to the
        exact test), 2 / 3 = left to the exact test (which says 0 / 1); raises if the grid is wrong"""
        abi = scn.to_abi()
        rc = self.lib.cwe_classify(C.byref(abi), C.c_double(x), C.c_double(y), C.c_double(th), C.c_int(agent))
        if rc < 0:
            raise AssertionError("grid classification disagrees with the exact test at (%r, %r, %r): %d" % (x, y, th, rc))
        return rc

    def check_terminal(self, scn, states, flags, horizon):
        abi = scn.to_abi()
        states = np.ascontiguousarray(states, np.float64)
        flags = np.ascontiguousarray(flags, np.uint8)
        return bool(self._check(self.lib.cwe_check_terminal(C.byref(abi), _p(states, C.c_double), _p(flags, C.c_uint8),
                                                            C.c_int(horizon))))

    def replay(self, scn, states0, flags0, horizon0, actions, q0=None):
        """-> brm_cw_trace-shaped dict with a batch dimension of 1."""
        abi = scn.to_abi()
        A, M, V = len(scn.agents), scn.n_mcts_agents, scn.params.REWARD_VECTOR_SIZE
        R = sum((A - 1) if r.IsAgentSpecific() else 1 for r in scn.rules)
        actions = np.ascontiguousarray(actions, np.uint8)
        T = actions.shape[0]
        states0 = np.ascontiguousarray(states0, np.float64)
        flags0 = np.ascontiguousarray(flags0, np.uint8)
        q0a = None if q0 is None else np.ascontiguousarray(q0, np.uint8)
        tr = dict(agents_out=np.zeros((1, T, A, 4)), flags_out=np.zeros((1, T, A), np.uint8),
                  terminal_out=np.zeros((1, T), np.uint8), q_out=np.zeros((1, T, M, R), np.uint8),
                  viol_out=np.zeros((1, T, M, R), np.uint32), rewards_out=np.zeros((1, T, M, V)),
                  labels_out=np.zeros((1, T, M, 4), np.uint32), frenet_out=np.zeros((1, T, A, 3)),
                  term_reward_out=np.zeros((1, M, V)), steps_out=np.zeros(1, np.int32))
        n = self._check(self.lib.cwe_replay(
            C.byref(abi), _p(states0, C.c_double), _p(flags0, C.c_uint8), C.c_int(horizon0), _p(q0a, C.c_uint8),
            _p(actions, C.c_uint8), C.c_int(T), _p(tr["agents_out"], C.c_double), _p(tr["flags_out"], C.c_uint8),
            _p(tr["terminal_out"], C.c_uint8), _p(tr["q_out"], C.c_uint8), _p(tr["viol_out"], C.c_uint32),
            _p(tr["rewards_out"], C.c_double), _p(tr["labels_out"], C.c_uint32), _p(tr["frenet_out"], C.c_double),
            _p(tr["term_reward_out"], C.c_double)))
        tr["steps_out"][0] = n
        return tr

    def rollout_detail(self, scn, states0, flags0, horizon0, seed, rollout_begin, n, max_steps, gamma,
                       first_discount_exp=1, q0=None, leaf_terminal=False):
        """-> brm_cw_rollout_out-shaped dict (ro_* arrays and agent_steps)."""
        abi = scn.to_abi()
        A, M, V = len(scn.agents), scn.n_mcts_agents, scn.params.REWARD_VECTOR_SIZE
        R = sum((A - 1) if r.IsAgentSpecific() else 1 for r in scn.rules)
        states0 = np.ascontiguousarray(states0, np.float64)
        flags0 = np.ascontiguousarray(flags0, np.uint8)
        q0a = None if q0 is None else np.ascontiguousarray(q0, np.uint8)
        o = dict(ro_returns=np.zeros((n, M, V)), ro_viol=np.zeros((n, M, R), np.uint32),
                 ro_q=np.zeros((n, M, R), np.uint8), ro_agents=np.zeros((n, A, 4)),
                 ro_flags=np.zeros((n, A), np.uint8), ro_steps=np.zeros(n, np.int32))
        steps = C.c_uint64(0)
        self._check(self.lib.cwe_rollout_detail(
            C.byref(abi), _p(states0, C.c_double), _p(flags0, C.c_uint8), C.c_int(horizon0), _p(q0a, C.c_uint8),
            C.c_uint64(seed), C.c_uint64(rollout_begin), C.c_uint64(n), C.c_int(max_steps), C.c_double(gamma),
            C.c_int(first_discount_exp), C.c_int(1 if leaf_terminal else 0), _p(o["ro_returns"], C.c_double),
            _p(o["ro_viol"], C.c_uint32), _p(o["ro_q"], C.c_uint8), _p(o["ro_agents"], C.c_double),
            _p(o["ro_flags"], C.c_uint8), _p(o["ro_steps"], C.c_int32), C.byref(steps)))
        o["agent_steps"] = int(steps.value)
        return o

    def evaluate_rules(self, scn, states, flags, horizon, q):
        abi = scn.to_abi()
        A, M, V = len(scn.agents), scn.n_mcts_agents, scn.params.REWARD_VECTOR_SIZE
        R = sum((A - 1) if r.IsAgentSpecific() else 1 for r in scn.rules)
        states = np.ascontiguousarray(states, np.float64)
        flags = np.ascontiguousarray(flags, np.uint8)
        q = np.ascontiguousarray(q, np.uint8)
        o = dict(q=np.zeros((M, R), np.uint8), viol=np.zeros((M, R), np.uint32), rewards=np.zeros((M, V)),
                 labels=np.zeros((M, 4), np.uint32))
        self._check(self.lib.cwe_evaluate_rules(
            C.byref(abi), _p(states, C.c_double), _p(flags, C.c_uint8), C.c_int(horizon), _p(q, C.c_uint8),
            _p(o["q"], C.c_uint8), _p(o["viol"], C.c_uint32), _p(o["rewards"], C.c_double),
            _p(o["labels"], C.c_uint32)))
        return o
