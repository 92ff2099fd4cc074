import sys; sys.path.insert(0,'.')
import numpy as np
import planner_rules_mcts_b200 as brm
from oracle import cw as ocw
from planner_rules_mcts_b200 import scenarios
from tests.cw_common import oracle_config, GUARD
np.set_printoptions(precision=7, suppress=True, linewidth=200)
orc = ocw.CwOracle()
scn = scenarios.merge_scenario()
p = scn.params
p.ACCELERATION_WEIGHT, p.RADIAL_ACCELERATION_WEIGHT, p.LANE_CENTER_WEIGHT = -0.5, -1.0, -2.0
p.GOAL_REWARD, p.COLLISION_WEIGHT = 100.0, -50.0
eng = brm.RulesMctsEngine([scn], device=0)
rng = np.random.default_rng(11)
B, T = 96, 22
acts = rng.integers(0, 3, size=(B, T, 4)).astype(np.uint8)
A=4
states = np.broadcast_to(scn.initial_state, (B, A, 4))
batch = eng.make_batch(states)
tr = eng.replay(batch, acts)
cfg = oracle_config(scn)
fl = np.full(A,3,np.uint8)
nbad=0
for b in range(B):
    o = orc.replay(cfg, states[b], fl, 20, acts[b])
    for t in range(o['steps']):
        if o['margin'][t] < GUARD: break
        d = np.abs(tr['rewards_out'][b,t].astype(np.float64) - o['rewards'][t])
        if (d > 1e-3 + 1e-5*np.abs(o['rewards'][t])).any():
            nbad+=1
            if nbad<4:
                print('b',b,'t',t,'\n gpu',tr['rewards_out'][b,t],'\n ref',o['rewards'][t],'\n gpu st',tr['agents_out'][b,t],'\n ref st',o['states'][t], '\n gpu fr', tr['frenet_out'][b,t], '\n ref fr', o['frenet'][t], '\n flags', tr['flags_out'][b,t], o['flags'][t], 'labels', tr['labels_out'][b,t].tolist(), o['labels'][t].tolist())
print('nbad', nbad)
# error statistics over all compared steps
mr=0.0; ms=0.0; mrel=0.0
for b in range(B):
    o = orc.replay(cfg, states[b], fl, 20, acts[b])
    for t in range(o['steps']):
        if o['margin'][t] < GUARD: break
        mr=max(mr, np.abs(tr['rewards_out'][b,t].astype(np.float64) - o['rewards'][t]).max())
        d=np.abs(tr['agents_out'][b,t].astype(np.float64) - o['states'][t])
        ms=max(ms, d.max()); mrel=max(mrel,(d/np.maximum(np.abs(o['states'][t]),1.0)).max())
print("max abs reward err", mr, "max abs state err", ms, "max rel state err", mrel)
