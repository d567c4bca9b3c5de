"""Small forward + backward + observation run for compute-sanitizer (memcheck / racecheck):
    compute-sanitizer --tool memcheck python tools/sanitize.py
Ragged sizes on purpose (N not a multiple of 32, partial last tiles)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from dyn_res_pile_manip_b200 import PlannerGD, PropNetDiffDenModel, ops, synthetic, observation
from dyn_res_pile_manip_b200.rewards import config_reward_ptcl

cfg, env = synthetic.default_config(), synthetic.FakeEnv()
torch.manual_seed(0)
model = PropNetDiffDenModel(cfg, True).cuda()
planner = PlannerGD(cfg, env)
goal = torch.from_numpy(synthetic.make_goal("bar")).cuda()
for mode in (0, 1, 2):
    ops.set_tensor_cores(mode)
    for N in (37, 130):
        planner.particle_num = N
        st, dn = synthetic.make_pile_batch(1, N, seed=1)
        acts = torch.tensor(synthetic.random_actions(5, 2, seed=2), device="cuda", requires_grad=True)
        out = planner.ptcl_model_rollout(torch.tensor(st).cuda(), torch.tensor(dn).cuda(), torch.zeros(1, N, device="cuda"),
                                         model, acts)
        pred = out["model_rollout"]["state_pred"]
        coor = planner.goal_coordinates(goal.cpu().numpy(), torch.device("cuda"))
        r = config_reward_ptcl(pred[:, -1], goal, env.get_cam_params(), coor)
        r.sum().backward()
        print("mode", mode, N, float(r.detach().mean()), float(acts.grad.abs().sum()))
st, _ = synthetic.make_pile_batch(1, 60, seed=3)
obs = synthetic.render_observation(st[0], env)
p, rad = observation.obs2ptcl_fixed_num_batch(obs, 37, 3, env.get_cam_params(), env.global_scale, seed=0)
print("obs", p.shape, float(rad.mean()))
torch.cuda.synchronize()
