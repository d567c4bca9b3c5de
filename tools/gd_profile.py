import cProfile, pstats, os, sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from dyn_res_pile_manip_b200 import PlannerGD, PropNetDiffDenModel, synthetic
cfg, env = synthetic.default_config(), synthetic.FakeEnv()
torch.manual_seed(0)
model = PropNetDiffDenModel(cfg, True).cuda()
planner = PlannerGD(cfg, env)
goal = synthetic.make_goal("bar")
st3, dn3 = synthetic.make_pile_batch(30, 100, seed=0)
act3 = synthetic.random_actions(50, 1, seed=9).transpose(1, 0, 2).astype(np.float64)
def call():
    return planner.trajectory_optimization_ptcl_multi_traj(st3, dn3, np.zeros((30, 100), np.float32), goal, model, act3,
                                                           np.zeros(1), 50, 1, 200, None, None, time_lim=2000)
for _ in range(4): call()
pr = cProfile.Profile(); pr.enable()
for _ in range(10): res = call()
pr.disable()
ps = pstats.Stats(pr); ps.sort_stats('cumulative').print_stats(28)
print(res['times'])
