import sys, os
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from dyn_res_pile_manip_b200 import PlannerGD, PropNetDiffDenModel, ops, synthetic
mode = int(sys.argv[1])
S, N, T = 128, 300, 20
cfg, env = synthetic.default_config(), synthetic.FakeEnv()
torch.manual_seed(0)
model = PropNetDiffDenModel(cfg, True).cuda()
planner = PlannerGD(cfg, env); planner.particle_num = N
st, dn = synthetic.make_pile_batch(1, N, seed=0)
s0, dens, attr = torch.tensor(st).cuda(), torch.tensor(dn).cuda(), torch.zeros(1, N).cuda()
acts = torch.tensor(synthetic.random_actions(S, T, seed=1), device="cuda", requires_grad=True)
ops.set_tensor_cores(mode)
for it in range(3):
    out = planner.ptcl_model_rollout(s0, dens, attr, model, acts)
    out["model_rollout"]["state_pred"].sum().backward()
torch.cuda.synchronize()
