"""include/brm_state.hpp: the C++ StateInterface adapter compiles against the C ABI, refuses
to run without a GPU, and on a GPU reproduces what the Python mirror computes."""
import os
import subprocess

import numpy as np
import pytest

import planner_rules_mcts_b200 as brm
from planner_rules_mcts_b200 import _abi, scenarios
from planner_rules_mcts_b200.gridworld import GW_LABELS

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "planner-rules-mcts_b200")
EXE = os.path.join(ROOT, "build", "adapter_test")


@pytest.fixture(scope="module")
def exe(built):
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    src = os.path.join(ROOT, "tests", "cpp", "adapter_test.cpp")
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), src, "-o", EXE,
                    "-L", PKG, "-lbrm", "-Wl,-rpath," + PKG], check=True)
    return EXE


def run(exe, *args):
    r = subprocess.run([exe] + list(args), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    return r.returncode, r.stdout


def test_adapter_compiles_and_refuses_without_gpu(exe):
    import torch
    rc, out = run(exe, "nogpu")
    assert rc == 0, out
    if not torch.cuda.is_available():
        assert "error code -2" in out and "no CPU fallback" in out


def test_cpp_host_tree_logic_on_the_cpu(built, tmp_path):
    """include/brm_mcts.hpp with a toy state and a toy evaluator (no engine): visit counts, virtual
    loss, back-propagation, thresholded lexicographic order, batched evaluator calls, terminal root."""
    exe = str(tmp_path / "mcts_logic_test")
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cpp", "mcts_logic_test.cpp"), "-o", exe, "-L", PKG, "-lbrm",
                    "-Wl,-rpath," + PKG], check=True)
    rc, out = run(exe)
    assert rc == 0 and "expectations hold" in out and "FAIL" not in out, out


@pytest.mark.parametrize("src", ["mcts_test.cpp", "planner_test.cpp", "comm_test.cpp"])
def test_cpp_hosts_compile_and_link_against_the_abi(built, tmp_path, src):
    """The C++ hosts of include/brm_mcts.hpp / brm_state.hpp build with warnings as errors and link
    against libbrm.so -- on the CPU, so a broken header shows up before a GPU box is involved."""
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cpp", src), "-o", str(tmp_path / "a.out"), "-L", PKG, "-lbrm",
                    "-lpthread", "-Wl,-rpath," + PKG], check=True)


@pytest.mark.gpu
def test_adapter_gridworld_matches_python_mirror(exe, tmp_path):
    rules = brm.make_default_rules()
    arr = (_abi.Rule * len(rules))(*[r.to_abi(GW_LABELS) for r in rules])
    path = tmp_path / "rules.bin"
    path.write_bytes(bytes(arr))
    n = 512
    rc, out = run(exe, "gw", str(path), str(n))
    assert rc == 0, out
    eng = brm.GridWorldEngine(device=0)
    s = brm.GridWorldState(eng)
    rewards = []
    lines = out.splitlines()
    step = 0
    while not s.IsTerminal() and step < 20:
        s = s.Execute([1 if step < 2 else 0, 0, 0], rewards)
        s.ResetDepth()
        x = s.GetAgentStates()[:, 0]
        want = "step %d x %d %d %d r0 %.17g %.17g %.17g" % (step, x[0], x[1], x[2], *rewards[0])
        assert lines[step] == want
        step += 1
    assert lines[step].startswith("terminal 1 steps 8 num_actions 1 Ego: x=8")
    res = eng.rollout(eng.base_test_env_batch(2), n, seed=1000, max_steps=30, discount_factor=0.95)
    means = res["returns_sum"] / n
    got = [l for l in lines if l.startswith("leaf")]
    assert len(got) == 6
    for l in range(2):
        for a in range(3):
            assert got[l * 3 + a] == "leaf %d agent %d mean %.17g %.17g %.17g" % (l, a, *means[l, a])
    stats = np.zeros((3, 3, 1 + eng.V))
    first = np.array([[0, 1, 2], [2, 1, 0]], np.uint8)
    eng.rollout(eng.base_test_env_batch(2), n, seed=1000, max_steps=30, discount_factor=0.95, first_action=first,
                stats=stats)
    got = [l for l in lines if l.startswith("stats")]
    assert len(got) == 9
    for a in range(3):
        for k in range(3):
            assert got[a * 3 + k] == "stats agent %d action %d visits %.17g sums %.17g %.17g %.17g" % (a, k, *stats[a, k])


@pytest.mark.gpu
def test_adapter_continuous_reference_pins(exe, tmp_path):
    scn = scenarios.straight_road_test_world()
    path = tmp_path / "scn.bin"
    path.write_bytes(bytes(scn.to_abi()))
    rc, out = run(exe, "cw", str(path))
    assert rc == 0, out
    lines = out.splitlines()
    assert lines[0] == "terminal0 0 num_actions 3"
    assert lines[1] == "a0 terminal 0 total 0"            # test/single_agent_test.cpp:143
    assert lines[2] == "a1 terminal 1 total -800"          # :152
    assert lines[3] == "a2 terminal 1 total -1000 num_actions 1"  # :159
    assert lines[4] == "execute on terminal: error -4"     # BARK_EXPECT_TRUE(!IsTerminal()), :59
    # the collided agent is INVALID, so its rules are not evaluated any more (:177); no violation at the start
    assert lines[5] == "evaluate_rules on collided state 0, on the initial state 0"


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["collision", "pass", "livelock"])
def test_cpp_host_tree_reproduces_the_threshold_suite(built, tmp_path, case):
    """include/brm_mcts.hpp: the leaf-parallel host search in C++ (no Python in the loop) reproduces
    test/grid_world_threshold_test.cpp:40-71 on the device engine, 3000 iterations x <= 16 decisions."""
    exe = os.path.join(ROOT, "build", "mcts_test")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cpp", "mcts_test.cpp"), "-o", exe, "-L", PKG, "-lbrm",
                    "-Wl,-rpath," + PKG], check=True)
    rules = brm.make_default_rules()
    arr = (_abi.Rule * len(rules))(*[r.to_abi(GW_LABELS) for r in rules])
    path = tmp_path / "rules.bin"
    path.write_bytes(bytes(arr))
    rc, out = run(exe, str(path), case, "8", "16")
    assert rc == 0 and "expectations hold" in out, out


@pytest.mark.gpu
@pytest.mark.parametrize("n_others,expected", [(0, 1), (1, 4)])
def test_cpp_host_tree_plans_in_the_continuous_world(built, tmp_path, n_others, expected):
    """include/brm_mcts.hpp over brm::MvmctsState / brm::RulesMctsRolloutEvaluator: the decision of
    BehaviorRulesMcts::Plan from C++ -- free road: accelerate (test/single_agent_test.cpp:248), a slower
    agent 2 m in front: brake (:268); the same pins tests/test_planner_gpu.py holds for planner.py."""
    from planner_rules_mcts_b200.ltl import RuleMonitor
    exe = os.path.join(ROOT, "build", "planner_test")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cpp", "planner_test.cpp"), "-o", exe, "-L", PKG, "-lbrm",
                    "-Wl,-rpath," + PKG], check=True)
    prims = [(0.0, 0.0), (5.0, 0.0), (0.0, -1.0), (0.0, 1.0), (-3.0, 0.0)]   # :191-200
    scn = scenarios.straight_road_test_world(
        n_others=n_others, rel_distance=2.0 if n_others else 7.0, ego_velocity=5.0,
        velocity_difference=2.0 if n_others else 0.0, primitives=prims, goal=(7.5, 12.5, -4.25, 0.75),
        params=brm.RulesMctsStateParameters(GOAL_REWARD=100.0),
        rules=[RuleMonitor.MakeRule("G !collision_ego", -1000.0, 0)])
    (tmp_path / "scn.bin").write_bytes(bytes(scn.to_abi()))
    (tmp_path / "state.bin").write_bytes(np.ascontiguousarray(scn.initial_state, np.float64).tobytes())
    rc, out = run(exe, str(tmp_path / "scn.bin"), str(tmp_path / "state.bin"), str(expected), "1000", "8", "16")
    assert rc == 0 and "expectations hold" in out, out
