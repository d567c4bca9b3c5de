"""Generate tests/golden/golden_big_v1.npz by RUNNING THE REAL REFERENCE at BASELINE.json's particle counts
(build container only; CPU, a few minutes):

    python tests/golden/make_golden_big.py

  F  one step, 8 state variants x 300 particles (config 3's largest pile): s_delta, s_pred, relation set
  G  rollout 2 state variants x 2 samples x 100 particles x T=10 (config 2's pile and horizon): states at every
     step, rewards of every step, d(sum -reward)/d(actions) through the reference's autograd, relation sets per step
  H  rollout 1 x 2 samples x 300 particles x T=20 (config 3 / 4's pile and horizon): the same quantities
  J, K  the same two rollouts with the predictor's output layer scaled by TAME = 0.02: the random-init network
     translates the whole pile by ~0.25 per step (it leaves the camera's view after a few steps and crosses z = 0
     around step 8), so G/H pin the arithmetic in that regime while J/K keep the pile in place, the pusher matters
     and the action gradient is dense -- the regime a trained checkpoint is in

Weights are the seed-0 reference initialisation already stored in golden_v1.npz (keys w/...).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import ref_harness  # noqa: E402
from make_golden import coo, ref_adjacency  # noqa: E402
from dyn_res_pile_manip_b200 import synthetic  # noqa: E402


TAME = 0.02


def rollout_case(ref, model, planner, G, tag, n_batch, n_sample, N, T, seed, goal_kind, tame=1.0):
    last = model.model.particle_predictor.linear_1
    with torch.no_grad():
        last.weight.mul_(tame)
        last.bias.mul_(tame)
    try:
        _rollout_case(ref, model, planner, G, tag, n_batch, n_sample, N, T, seed, goal_kind)
    finally:
        with torch.no_grad():
            last.weight.div_(tame)
            last.bias.div_(tame)
    G[tag + "/tame"] = np.array(tame, dtype=np.float32)


def _rollout_case(ref, model, planner, G, tag, n_batch, n_sample, N, T, seed, goal_kind):
    st, dn = synthetic.make_pile_batch(n_batch, N, seed=seed)
    acts = synthetic.random_actions(n_sample * n_batch, T, seed=seed, lim=3.0)
    goal = synthetic.make_goal(goal_kind)
    goal_t = torch.from_numpy(goal)
    coords = torch.flip((goal_t < 0.5).nonzero(), dims=(1,)).float().numpy()
    goal_coor, _ = ref["utils"].fps_np(coords, min(5 * N, coords.shape[0]), 0)
    planner.particle_num = N
    act_t = torch.tensor(acts, requires_grad=True)
    with ref_harness.cpu_cuda_shims():
        out = planner.ptcl_model_rollout(torch.from_numpy(st), torch.from_numpy(dn), torch.zeros(n_batch, N), model, act_t)
    pred = out["model_rollout"]["state_pred"]
    obs = pred.reshape(n_sample * n_batch, 1, T, N, 3).permute(0, 2, 1, 3, 4)
    reward, next_r = planner.ptcl_evaluate_traj(obs, goal_t, torch.from_numpy(goal_coor).float())
    torch.sum(-reward).backward()
    G[tag + "/dims"] = np.array([n_batch, n_sample, N, T])
    G[tag + "/s0"], G[tag + "/dens"], G[tag + "/acts"] = st, dn, acts
    G[tag + "/goal_kind"], G[tag + "/goal_coor"] = np.array(goal_kind), goal_coor.astype(np.float32)
    G[tag + "/state_pred"] = pred.detach().numpy()
    G[tag + "/reward"] = reward.detach().numpy()
    G[tag + "/next_r"] = next_r.detach().numpy()
    G[tag + "/act_grad"] = act_t.grad.numpy()
    s = torch.from_numpy(st).repeat(n_sample, 1, 1)
    moved = 0
    for t in range(T):
        with torch.no_grad():
            sdel = planner.gen_s_delta(s, act_t.detach()[:, t])
            moved += int((sdel.abs().sum(-1) > 0).sum())
            G[tag + "/rel%d" % t] = coo(ref_adjacency(model, s, sdel))
            s = pred.detach()[:, t]
    print(tag, "done; pushed particle-steps:", moved, "reward", reward.detach().numpy().ravel())


def main():
    ref = ref_harness.load_reference()
    torch.set_num_threads(8)
    cfg, env = synthetic.default_config(), synthetic.FakeEnv()
    G = {}
    torch.manual_seed(0)
    model = ref["gnn_dyn"].PropNetDiffDenModel(cfg, False)
    base = np.load(os.path.join(HERE, "golden_v1.npz"))
    for k, v in model.state_dict().items():
        assert np.array_equal(base["w/" + k], v.numpy()), k       # same seed-0 weights as golden_v1
    planner = ref["planners"].PlannerGD(cfg, env)

    # ---- F: one step at 300 particles --------------------------------------------------------------------
    states, dens = synthetic.make_pile_batch(8, 300, seed=21)
    act = synthetic.random_actions(8, 1, seed=21, lim=3.0)[:, 0]
    planner.particle_num = 300
    s_cur = torch.from_numpy(states)
    s_delta = planner.gen_s_delta(s_cur, torch.from_numpy(act))
    with torch.no_grad():
        s_pred = model.predict_one_step(torch.zeros(8, 300), s_cur, s_delta, torch.from_numpy(dens))
    G["F/s_cur"], G["F/act"], G["F/dens"] = states, act, dens
    G["F/s_delta"], G["F/s_pred"] = s_delta.numpy(), s_pred.numpy()
    G["F/rel"] = coo(ref_adjacency(model, s_cur, s_delta))
    print("F done; pushed particles:", int((s_delta.abs().sum(-1) > 0).sum()))

    rollout_case(ref, model, planner, G, "G", 2, 2, 100, 10, 22, "tee")
    rollout_case(ref, model, planner, G, "H", 1, 2, 300, 20, 23, "bar")
    rollout_case(ref, model, planner, G, "J", 2, 2, 100, 10, 22, "tee", tame=TAME)
    rollout_case(ref, model, planner, G, "K", 1, 2, 300, 20, 23, "bar", tame=TAME)

    path = os.path.join(HERE, "golden_big_v1.npz")
    np.savez_compressed(path, **G)
    print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1e3), len(G), "arrays")


if __name__ == "__main__":
    main()
