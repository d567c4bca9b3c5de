"""Where the reference MPC entry point's time goes (measurement aid): wall vs. device time of one
trajectory_optimization_ptcl_multi_traj call with the shipped configuration (50 traj x 30 variants x 100 particles)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dyn_res_pile_manip_b200 import PlannerGD, PropNetDiffDenModel, synthetic
cfg, env = synthetic.default_config(), synthetic.FakeEnv()
torch.manual_seed(0)
model = PropNetDiffDenModel(cfg, True).cuda()
planner = PlannerGD(cfg, env)
goal = synthetic.make_goal("bar")
st, dn = synthetic.make_pile_batch(30, 100, seed=0)
act = synthetic.random_actions(50, 1, seed=9).transpose(1, 0, 2).astype(np.float64)
for i in range(5):
    t = time.perf_counter()
    res = planner.trajectory_optimization_ptcl_multi_traj(st, dn, np.zeros((30, 100), np.float32), goal, model, act,
                                                          np.zeros(1), 50, 1, 200, None, None, time_lim=2000)
    wall = (time.perf_counter() - t) * 1e3
    tm = res["times"]
    print("call %d: wall %.1f ms, iterations %d, device rollout %.1f + optim %.1f ms, total_time %.1f ms" %
          (i, wall, res["iter_num"] + 1, tm["rollout_time"], tm["optim_time"], tm["total_time"] * 1e3))
