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

# ---- round 2 paths: training step (weight gradients), real-robot pusher, graph-less GD planner loop with the device-side
# bookkeeping kernels, multi-scene planning, resolution regressor
ops.set_tensor_cores(2)
torch.manual_seed(0)
tm = PropNetDiffDenModel(cfg, True).cuda()
st, dn = synthetic.make_pile_batch(3, 45, seed=2)
nums = torch.tensor([45, 31, 20])
s = torch.tensor(st).cuda()
sd = torch.tensor((np.random.RandomState(0).normal(0, 0.01, st.shape)).astype(np.float32)).cuda()
out = tm.predict_one_step(torch.zeros(3, 45).cuda(), s, sd, torch.tensor(dn).cuda(), nums)
out2 = tm.predict_one_step(torch.zeros(3, 45).cuda(), out, sd, torch.tensor(dn).cuda(), nums)
(out2 ** 2).mean().backward()
assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in tm.parameters())
planner.use_graph = False
st, dn = synthetic.make_pile_batch(3, 37, seed=3)
act0 = synthetic.random_actions(4, 2, seed=3).transpose(1, 0, 2).astype(np.float64)
res = planner.trajectory_optimization_ptcl_multi_scene([st, st + 0.001], [dn, dn], [np.zeros((3, 37), np.float32)] * 2,
                                                       synthetic.make_goal("bar"), model, act0, np.zeros(2), 4, 2, 2)
assert len(res) == 2 and np.isfinite(res[0]["action_sequence"]).all()
from dyn_res_pile_manip_b200 import MPCResRgrNoPool
rg = MPCResRgrNoPool({"train_res_cls": {"state_h": 224, "state_w": 224, "res_dim": 6}})
y = rg.forward(torch.rand(2, 6, 224, 224).cuda())
assert torch.isfinite(y).all()
# general-width engine (csrc/general.cu): width 150 (padded to 192, streamed weight blocks) and 96: training step with
# weight gradients on a padded batch, planner rollout with action gradients
import copy
for nf in (150, 96):
    cfg_w = copy.deepcopy(cfg)
    cfg_w['train']['particle']['nf_effect'] = nf
    torch.manual_seed(0)
    wm = PropNetDiffDenModel(cfg_w, True).cuda()
    st, dn = synthetic.make_pile_batch(3, 45, seed=2)
    s = torch.tensor(st).cuda()
    out = wm.predict_one_step(torch.zeros(3, 45).cuda(), s, sd, torch.tensor(dn).cuda(), torch.tensor([45, 31, 20]))
    (out ** 2).mean().backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in wm.parameters())
    planner.particle_num = 37
    st, dn = synthetic.make_pile_batch(1, 37, seed=1)
    acts = torch.tensor(synthetic.random_actions(5, 2, seed=2), device="cuda", requires_grad=True)
    pred = planner.ptcl_model_rollout(torch.tensor(st).cuda(), torch.tensor(dn).cuda(), torch.zeros(1, 37, device="cuda"),
                                      wm, acts)["model_rollout"]["state_pred"]
    pred.sum().backward()
    assert torch.isfinite(acts.grad).all()
    # inference (no gradient): hoisted relation propagator, wide relation-side layers on tcgen05 (csrc/general_tc.cu;
    # 4500 relation slots, ragged)
    st, dn = synthetic.make_pile_batch(10, 45, seed=3)
    with torch.no_grad():
        out = wm.predict_one_step(torch.zeros(10, 45).cuda(), torch.tensor(st).cuda(), 0.01 * torch.randn(10, 45, 3).cuda(),
                                  torch.tensor(dn).cuda(), torch.tensor([45, 31, 20, 45, 45, 2, 45, 44, 45, 45]))
    assert torch.isfinite(out).all()
torch.cuda.synchronize()
print("round-2 paths ok")
