import sys, time
import numpy as np
sys.path.insert(0, '/root/repo')
import planner_rules_mcts_b200 as brm
from planner_rules_mcts_b200 import gridworld_env as ge
BIG = np.finfo(np.float64).max
eng = None
for iters, rpe in [(3000, 16), (3000, 1), (3000, 4), (1000, 64)]:
    ok = 0
    for seed in range(1000, 1008):
        env = ge.CrossingTestEnv((-0.37, -0.37, BIG), engine=eng, rollouts_per_expansion=rpe, seed=seed)
        eng = env.engine
        t0 = time.time()
        res = ge.TestRunner(env).RunTest(iters, 16)
        h = np.array(env.state_history)[-1]
        good = h[0] > 8 and h[1] > 8 and h[2] > 8 and h[0] < h[2] < h[1]
        ok += good
        print(iters, rpe, seed, "final", h, good, "%.1fs" % (time.time() - t0), flush=True)
    print("==", iters, rpe, ok, "of 8", flush=True)
