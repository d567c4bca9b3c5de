"""p50 of the reference's MPC entry (50 traj x 30 variants x 100 particles, T=1, 27 Adam iterations) and of one config-4
iteration, for A/B runs of two libraries:  [PILE_GNN_LIB=...] python tools/gd_time.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from dyn_res_pile_manip_b200 import PlannerGD, PropNetDiffDenModel, synthetic

cfg, env = synthetic.default_config(), synthetic.FakeEnv()
torch.manual_seed(0)
model = PropNetDiffDenModel(cfg, True).cuda()
planner = PlannerGD(cfg, env)
goal = synthetic.make_goal("bar")
st3, dn3 = synthetic.make_pile_batch(30, 100, seed=0)
act3 = synthetic.random_actions(50, 1, seed=9).transpose(1, 0, 2).astype(np.float64)
gd = []
for i in range(12):
    t_a = time.perf_counter()
    res = planner.trajectory_optimization_ptcl_multi_traj(st3, dn3, np.zeros((30, 100), np.float32), goal, model, act3,
                                                          np.zeros(1), 50, 1, 200, None, None, time_lim=2000)
    gd.append((time.perf_counter() - t_a) * 1e3)
gd = sorted(gd[2:])
st4, dn4 = synthetic.make_pile_batch(1, 300, seed=0)
act4 = synthetic.random_actions(128, 20, seed=4).transpose(1, 0, 2).astype(np.float64)
best = None
for rep in range(3):
    r4 = planner.trajectory_optimization_ptcl_multi_traj(st4, dn4, np.zeros((1, 300), np.float32), goal, model, act4,
                                                         np.zeros(20), 128, 20, 6, None, None, rollout_best_action_sequence=False)
    cur = (r4["times"]["rollout_time"] / 6, r4["times"]["optim_time"] / 6)
    best = cur if best is None or sum(cur) < sum(best) else best
print("[%s] MPC entry p50 %.2f ms (loop %.2f ms, reward %.6f) | cfg4 fwd %.2f + bwd %.2f ms | action sum %.6f" % (
    os.path.basename(os.path.dirname(os.environ.get("PILE_GNN_LIB", "./default"))), gd[len(gd) // 2],
    res["times"]["rollout_time"] + res["times"]["optim_time"], float(res["reward"].sum()), best[0], best[1],
    float(np.abs(res["action_full"]).sum())))
