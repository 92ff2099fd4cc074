import sys; sys.path.insert(0,'.')
import numpy as np
import planner_rules_mcts_b200 as brm
from planner_rules_mcts_b200 import planner
from tests.test_planner_gpu import suite_world
np.set_printoptions(precision=2, suppress=True, linewidth=200)
scn = suite_world(1, 2.0, 2.0, goal=(30.0, 60.0, -4.25, 0.75))
pl = planner.BehaviorRulesMcts(scn, planner.PlannerParameters(MaxNumIterations=1000, RolloutsPerExpansion=32))
eng = pl.engine
cur = eng.initial_batch([0])
for t in range(40):
    if cur.terminal[0]: break
    a = pl.Plan(cur)
    act = np.full((eng.M,1),0,np.uint8); act[0,0]=a
    cur,_ = eng.execute(cur, act)
    if not cur.terminal[0] or cur.horizon[0]==0:
        cur.horizon[:] = 30; eng.check_terminal(cur)
    print(t, 'a', a, 'ego', cur.agents[0,0], 'other', cur.agents[1,0], 'term', cur.terminal[0], 'flags', cur.flags[:,0], 'N', pl.last_stats[0,:,0].astype(int))
