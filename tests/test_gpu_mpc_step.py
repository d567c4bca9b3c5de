"""Planner-side half of a closed-loop MPC step (dyn_res_pile_manip_b200.mpc) on the GPU: observation ->
particles -> plan -> warm start, with the model's own prediction standing in for the simulator."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_two_closed_loop_steps():
    from dyn_res_pile_manip_b200 import PlannerGD, PropNetDiffDenModel, synthetic
    from dyn_res_pile_manip_b200.mpc import MPCStep
    from dyn_res_pile_manip_b200.rewards import config_reward_ptcl
    cfg, env = synthetic.default_config(), synthetic.FakeEnv()
    torch.manual_seed(0)
    model = PropNetDiffDenModel(cfg, True).cuda()
    planner = PlannerGD(cfg, env)
    N, n_sample, n_look = 100, 6, 1
    st, _ = synthetic.make_pile_batch(1, 300, seed=4)
    goal = synthetic.make_goal("bar")
    step = MPCStep(planner, model, env, batch_size=30)
    init = synthetic.random_actions(n_sample, 1, seed=4).transpose(1, 0, 2).astype(np.float64)      # [1, traj, 4]
    labels = np.zeros(1)
    state = st[0]
    lo, hi = planner.action_box(0)
    for it in range(2):
        obs = synthetic.render_observation(state, env)
        out = step.plan(obs, goal, N, init, labels, n_sample, n_look, 6, None, None, seed=it)
        assert out["obs_cur"].shape == (30, N, 3) and np.isfinite(out["obs_cur"]).all()
        assert out["particle_den"].shape == (30,) and (out["particle_den"] > 0).all()
        a = out["action"]
        assert a.shape == (4,) and (a >= lo - 1e-6).all() and (a <= hi + 1e-6).all()
        # the reported reward is config_reward_ptcl of the first re-sampling
        planner.particle_num = N
        coor = planner.goal_coordinates(goal, torch.device("cuda"))
        r = config_reward_ptcl(torch.from_numpy(out["obs_cur"]).float().cuda(), torch.from_numpy(goal).float().cuda(),
                               env.get_cam_params(), coor, normalize=True)[0].item()
        assert out["reward"] == r
        # warm start of a length-1 sequence is the sequence itself (flex_env.py:1112)
        assert out["action_seq_mpc_init"] is init and out["action_label_seq_mpc_init"] is labels
        # stand-in for the simulator: the model's own prediction for the chosen push, at full resolution
        full = torch.from_numpy(state[None]).cuda()
        dens = torch.tensor([float(out["particle_den"][0])], device="cuda")
        planner.particle_num = 300
        roll = planner.ptcl_model_rollout(full, dens, torch.zeros(1, 300, device="cuda"), model,
                                          torch.from_numpy(a[None, None]).float().cuda())
        state = roll["model_rollout"]["state_pred"][0, 0].detach().cpu().numpy()
        assert np.isfinite(state).all()
