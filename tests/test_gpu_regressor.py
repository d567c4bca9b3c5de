"""Resolution regressor on the device (dyn_res_pile_manip_b200.regressor, csrc/rgr.cu) against the output of the real
reference (tests/golden/golden_rgr_v1.npz) and against torch's own CPU evaluation of the same module; then the
dynamic-resolution MPC loop: the regressor picks the particle count of every step."""
import os

import numpy as np
import pytest
import torch

from dyn_res_pile_manip_b200 import MPCResRgrNoPool, PlannerGD, PropNetDiffDenModel, synthetic
from dyn_res_pile_manip_b200.mpc import MPCStep

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
CFG = {"train_res_cls": {"state_h": 224, "state_w": 224, "res_dim": 6}}


@pytest.fixture(scope="module")
def golden_rgr():
    return np.load(os.path.join(HERE, "golden", "golden_rgr_v1.npz"))


@pytest.fixture(scope="module")
def net(golden_rgr):
    torch.manual_seed(int(golden_rgr["seed"]))          # the reference's seeded initialisation
    return MPCResRgrNoPool(CFG)


def test_forward_and_infer_param_vs_reference(golden_rgr, net):
    g = golden_rgr
    y = net.forward(torch.from_numpy(g["x"][None]).cuda())
    np.testing.assert_allclose(y.cpu().numpy(), g["y"], rtol=2e-4, atol=2e-6)
    fg, goal = g["fg"].astype(np.float32), g["goal"].astype(np.float32)
    assert net.infer_param(fg, goal) == int(g["particle_num"])
    assert net.infer_param(fg, goal) == int(g["particle_num"])          # graph replay


def test_batched_forward_vs_torch_cpu(net):
    rng = np.random.RandomState(0)
    x = torch.from_numpy(rng.uniform(0, 1, (5, 6, 224, 224)).astype(np.float32))
    with torch.no_grad():
        ref = net.model(x)                                      # torch's own kernels on the CPU, same weights
    got = net.forward(x.cuda()).cpu()
    np.testing.assert_allclose(got.numpy(), ref.numpy(), rtol=2e-4, atol=2e-6)
    again = net.forward(x.cuda()).cpu()
    assert torch.equal(got, again)                              # fixed summation order: deterministic
    # a changed parameter is picked up (re-pack keyed on the parameters' versions)
    with torch.no_grad():
        net.model[19].bias.add_(3.0)
    try:
        np.testing.assert_allclose(net.forward(x.cuda()).cpu().numpy(), ref.numpy() + 3.0, rtol=2e-4, atol=1e-5)
    finally:
        with torch.no_grad():
            net.model[19].bias.sub_(3.0)
    with pytest.raises(Exception):
        net.forward(torch.zeros(1, 6, 200, 200).cuda())         # does not reduce to 7 x 7


def test_dynamic_resolution_closed_loop(net):
    """MPCStep.plan without a particle count: three closed-loop steps, the regressor selects N from the current
    foreground mask and the goal every step (flex_env.py:981-998, 1080-1090); the captured planner loops follow."""
    cfg, env = synthetic.default_config(), synthetic.FakeEnv()
    torch.manual_seed(0)
    model = PropNetDiffDenModel(cfg, True).cuda()
    planner = PlannerGD(cfg, env)
    # make the (randomly initialised) regressor's output a usable, image-dependent particle count
    keep = (net.model[19].weight.detach().clone(), net.model[19].bias.detach().clone())
    with torch.no_grad():                       # random init: w.h = -0.01664 +- 5e-5 over images -> N = 100 +- 25
        net.model[19].weight.mul_(5e5)
        net.model[19].bias.fill_(8421.6)
    try:
        step = MPCStep(planner, model, env, batch_size=8, res_rgr=net)
        goal = synthetic.make_goal("bar")
        st, _ = synthetic.make_pile_batch(1, 300, seed=4)
        state = st[0]
        init = synthetic.random_actions(5, 1, seed=4).transpose(1, 0, 2).astype(np.float64)
        seen = []
        for it in range(3):
            obs = synthetic.render_observation(state, env)
            fg = (obs[..., -1] / env.global_scale < 0.599 / 0.8).astype(np.float32)
            from dyn_res_pile_manip_b200.regressor import regressor_input
            with torch.no_grad():
                want = int(net.model(torch.from_numpy(regressor_input(fg, (goal < 0.5).astype(np.float32), 224, 224)[None])).item())
            out = step.plan(obs, goal, None, init, np.zeros(1), n_update_iter=4, seed=it)
            N = out["particle_num"]
            assert abs(N - want) <= 1 and N > 10, (N, want)    # int() of two fp32 evaluations may straddle an integer
            assert out["obs_cur"].shape == (8, N, 3)
            assert out["traj_opt_out"]["observation_sequence"].shape == (1, N, 3)
            seen.append(N)
            # stand-in simulator: push the full-resolution pile with the chosen action
            planner.particle_num = 300
            roll = planner.ptcl_model_rollout(torch.from_numpy(state[None]).cuda(),
                                              torch.tensor([float(out["particle_den"][0])], device="cuda"),
                                              torch.zeros(1, 300, device="cuda"), model,
                                              torch.from_numpy(out["action"][None, None]).float().cuda())
            new = roll["model_rollout"]["state_pred"][0, 0].detach().cpu().numpy()
            state = np.concatenate([new[:, :2] * (0.85 ** (it + 1)), state[:, 2:]], axis=1)      # keep it in view, shrink
        print("particle counts picked:", seen)
        # snapping to buckets
        step_b = MPCStep(planner, model, env, batch_size=8, res_rgr=net, resolution_buckets=[50, 100, 200, 300])
        out = step_b.plan(synthetic.render_observation(st[0], env), goal, None, init, np.zeros(1), n_update_iter=2, seed=0)
        assert out["particle_num"] in (50, 100, 200, 300)
    finally:
        with torch.no_grad():
            net.model[19].weight.copy_(keep[0])
            net.model[19].bias.copy_(keep[1])
