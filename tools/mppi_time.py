"""p50 of PlannerGD.trajectory_optimization_mppi (config 2: 256 samples x 100 particles x T=10, 3 iterations, numpy in ->
numpy plan out) next to the device time of its three captured evaluations:  python tools/mppi_time.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from dyn_res_pile_manip_b200 import PlannerGD, PropNetDiffDenModel, synthetic

cfg, env = synthetic.default_config(), synthetic.FakeEnv()
torch.manual_seed(0)
model = PropNetDiffDenModel(cfg, True).cuda()
planner = PlannerGD(cfg, env)
S, N, T = 256, 100, 10
st, dn = synthetic.make_pile_batch(1, N, seed=0)
goal = synthetic.make_goal("disc")
mean = synthetic.random_actions(1, T, seed=3)[0]
ml = []
for i in range(45):
    t0 = time.perf_counter()
    out = planner.trajectory_optimization_mppi(st, dn, np.zeros((1, N), np.float32), goal, model, mean, n_sample=S, n_update_iter=3, seed=i)
    ml.append((time.perf_counter() - t0) * 1e3)
ml = sorted(ml[5:])
eng = next(iter(planner._mppi_engines.values()))
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(30):
    eng.evaluate()
b.record(); torch.cuda.synchronize()
print("mppi call p50 %.2f ms  p90 %.2f ms | one captured evaluation %.3f ms on the device | plan sum %.6f" %
      (ml[len(ml) // 2], ml[int(len(ml) * 0.9)], a.elapsed_time(b) / 30, float(np.sum(out['action_sequence']))))
