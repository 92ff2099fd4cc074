"""The GridWorld threshold suite over several seeds / settings (see
tests/test_gridworld_threshold_gpu.py): usage gw_threshold_sweep.py <iterations> <rollouts per expansion>."""
import sys, time
import numpy as np
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from tests import gridworld_env as ge
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
rpe = int(sys.argv[2]) if len(sys.argv) > 2 else 16
BIG = np.finfo(np.float64).max
eng = None
for stat in ("greedy",):
  for name, thr in [("collision", (-1.0, -1.0, BIG)), ("pass", (-0.37, -0.37, BIG)), ("livelock", (-0.1, -0.8, BIG))]:
    for seed in (1000, 1001, 1002):
        env = ge.CrossingTestEnv(thr, engine=eng, rollouts_per_expansion=rpe, seed=seed, statistic=stat)
        eng = env.engine
        t0 = time.time()
        res = ge.TestRunner(env).RunTest(iters, 16)
        h = np.array(env.state_history)
        print(name, seed, "final", h[-1], "steps", len(h) - 1, "res", str(res).replace("\t", " "), "%.1fs" % (time.time() - t0), flush=True)
        print("   actions", env.action_history, flush=True)
