"""A few model steps at a given batch for ncu:  python tools/ncu_probe.py <samples> [particles] [steps] [tape]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dyn_res_pile_manip_b200 import PlannerGD, PropNetDiffDenModel, ops, synthetic

S = int(sys.argv[1]); N = int(sys.argv[2]) if len(sys.argv) > 2 else 300; T = int(sys.argv[3]) if len(sys.argv) > 3 else 2
tape = len(sys.argv) > 4 and sys.argv[4] == "tape"
cfg, env = synthetic.default_config(), synthetic.FakeEnv()
torch.manual_seed(0)
model = PropNetDiffDenModel(cfg, True).cuda()
planner = PlannerGD(cfg, env); planner.particle_num = N
st, dn = synthetic.make_pile_batch(1, N, seed=0)
s0, dens, attr = torch.tensor(st).cuda(), torch.tensor(dn).cuda(), torch.zeros(1, N).cuda()
acts = torch.tensor(synthetic.random_actions(S, T, seed=1), device="cuda", requires_grad=tape)
out = planner.ptcl_model_rollout(s0, dens, attr, model, acts)
if tape:
    out["model_rollout"]["state_pred"].sum().backward()
torch.cuda.synchronize()
