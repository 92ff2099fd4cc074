"""Root-parallel merge behind the C ABI (brm_comm_init / brm_allreduce_root_stats, NCCL bound at
run time) and engines on several GPUs of one process."""
import os
import subprocess

import numpy as np
import pytest

import planner_rules_mcts_b200 as brm
from planner_rules_mcts_b200 import _abi, scenarios
from planner_rules_mcts_b200.gridworld import GW_LABELS

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "planner-rules-mcts_b200")


def n_gpus():
    import torch
    return torch.cuda.device_count()


def test_single_rank_communicator_is_the_identity(built):
    scn = scenarios.merge_scenario()
    eng = brm.RulesMctsEngine([scn])
    rng = np.random.default_rng(3)
    L, n, n_act = 16, 32, 3
    leaves = eng.initial_batch([0]).select(np.zeros(L, np.int64))
    first = rng.integers(0, n_act, size=(L, eng.M)).astype(np.uint8)
    want = np.zeros((eng.M, n_act, 1 + eng.V))
    eng.rollout(leaves, n, seed=11, max_steps=12, first_action=first, stats=want)
    assert eng.engine.comm_size() == 1
    eng.engine.comm_init(_abi.comm_unique_id(), 0, 1)
    assert eng.engine.comm_size() == 1
    got = np.zeros_like(want)
    eng.rollout(leaves, n, seed=11, max_steps=12, first_action=first, stats=got)
    assert np.array_equal(got, want)
    buf = want.copy()
    eng.engine.allreduce_root_stats(buf)
    assert np.array_equal(buf, want)
    eng.engine.comm_destroy()


def test_engines_on_two_gpus_of_one_process(built):
    """One handle = one GPU: the second engine of a process must get its own kernel attributes
    (opt-in shared memory is per device) and compute the same bits as the first."""
    if n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    scn = scenarios.merge_scenario()
    outs = []
    for dev in (0, 1):
        eng = brm.RulesMctsEngine([scn], device=dev)
        outs.append(eng.rollout(eng.initial_batch(), 2048, seed=5, max_steps=20, detail=True))
    for k in outs[0]:
        assert np.array_equal(outs[0][k], outs[1][k]), k


def test_cpp_engines_merge_root_statistics(built, tmp_path):
    """tests/cpp/comm_test.cpp: one engine per GPU in one C++ process, merged through the C ABI."""
    n = min(n_gpus(), 8)
    exe = os.path.join(ROOT, "build", "comm_test")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-pthread", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cpp", "comm_test.cpp"), "-o", exe, "-L", PKG, "-lbrm",
                    "-Wl,-rpath," + PKG], check=True)
    rules = brm.make_default_rules()
    arr = (_abi.Rule * len(rules))(*[r.to_abi(GW_LABELS) for r in rules])
    path = tmp_path / "rules.bin"
    path.write_bytes(bytes(arr))
    r = subprocess.run([exe, str(path), str(n), "256"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                       timeout=300)
    assert r.returncode == 0, r.stdout
    last = r.stdout.strip().splitlines()[-1]
    assert last.startswith("ranks %d mismatches 0 visits %d" % (n, n * 2 * 256 * 3)), r.stdout
    if n > 1:
        assert last.endswith("differ 1")      # the ranks did contribute different rollouts
