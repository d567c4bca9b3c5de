"""Generate tests/golden/golden_v1.npz by RUNNING THE REAL REFERENCE (build container only).

    python tests/golden/make_golden.py

Imports /root/reference through `ref_harness` (stubs for matplotlib/dgl/open3d), drives
`PropNetDiffDenModel.predict_one_step`, `PlannerGD.gen_s_delta / ptcl_model_rollout /
ptcl_evaluate_traj / sample_action_sequences / optimize_action` on seeded synthetic piles
and stores inputs + outputs.  The fixtures travel to the GPU box; the reference does not.
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
from dyn_res_pile_manip_b200 import synthetic  # noqa: E402


def coo(adj):
    """bool [B,N,N] -> int16 [E,3] rows (b, recv, send) in nonzero order."""
    return adj.nonzero().to(torch.int16).numpy()


def ref_adjacency(model, s_cur, s_delta, particle_nums=None):
    """Recover the relation set the reference builds by intercepting model.forward."""
    grabbed = {}
    inner = model.model
    orig = inner.forward

    def spy(a_cur, s, sd, Rr, Rs, dens, verbose=False):
        grabbed["Rr"], grabbed["Rs"] = Rr.clone(), Rs.clone()
        return orig(a_cur, s, sd, Rr, Rs, dens)

    inner.forward = spy
    try:
        B, N, _ = s_cur.shape
        out = model.predict_one_step(torch.zeros(B, N), s_cur, s_delta, torch.ones(B), particle_nums)
    finally:
        inner.forward = orig
    Rr, Rs = grabbed["Rr"], grabbed["Rs"]
    adj = torch.zeros(B, N, N, dtype=torch.bool)
    for b in range(B):
        live = Rr[b].sum(1) > 0
        r = Rr[b][live].argmax(1)
        s = Rs[b][live].argmax(1)
        adj[b, r, s] = True
    del out
    return adj


def main():
    ref = ref_harness.load_reference()
    torch.set_num_threads(4)
    cfg = synthetic.default_config()
    env = synthetic.FakeEnv()
    G = {}

    torch.manual_seed(0)
    model = ref["gnn_dyn"].PropNetDiffDenModel(cfg, False)
    sd = model.state_dict()
    for k, v in sd.items():
        G["w/" + k] = v.numpy().copy()
    planner = ref["planners"].PlannerGD(cfg, env)
    G["cam_extrinsic"] = env.get_cam_extrinsics()
    G["cam_params"] = np.array(env.get_cam_params(), dtype=np.float64)

    # ---- case A: one step, 4 state variants x 100 particles, pusher through the pile -------------
    states, dens = synthetic.make_pile_batch(4, 100, seed=0)
    act = np.array([[-3.0, 0.5, 3.0, -0.5], [0.5, -3.5, -0.5, 3.0], [-2.0, -2.0, 2.5, 2.0],
                    [3.0, 1.0, -3.0, 1.5]], dtype=np.float32)
    s_cur = torch.from_numpy(states)
    planner.particle_num = 100
    s_delta = planner.gen_s_delta(s_cur, torch.from_numpy(act))
    a_cur = torch.zeros(4, 100)
    with torch.no_grad():
        s_pred = model.predict_one_step(a_cur, s_cur, s_delta, torch.from_numpy(dens))
    G["A/s_cur"], G["A/act"], G["A/dens"] = states, act, dens
    G["A/s_delta"] = s_delta.numpy()
    G["A/s_pred"] = s_pred.numpy()
    G["A/rel"] = coo(ref_adjacency(model, s_cur, s_delta))

    # ---- case B: non-zero attrs, tiny N (< 10 -> k = N), and the particle_nums padding mask -------
    states7, dens7 = synthetic.make_pile_batch(3, 7, seed=1)
    rng = np.random.RandomState(5)
    a7 = rng.uniform(0, 1, size=(3, 7)).astype(np.float32)
    sd7 = (rng.normal(0, 0.01, size=(3, 7, 3))).astype(np.float32)
    with torch.no_grad():
        p7 = model.predict_one_step(torch.from_numpy(a7), torch.from_numpy(states7), torch.from_numpy(sd7),
                                    torch.from_numpy(dens7))
    G["B/s_cur"], G["B/a_cur"], G["B/s_delta"], G["B/dens"], G["B/s_pred"] = states7, a7, sd7, dens7, p7.numpy()
    G["B/rel"] = coo(ref_adjacency(model, torch.from_numpy(states7), torch.from_numpy(sd7)))

    states40, dens40 = synthetic.make_pile_batch(3, 40, seed=2)
    sd40 = (rng.normal(0, 0.01, size=(3, 40, 3))).astype(np.float32)
    nums = torch.tensor([40, 25, 33])
    with torch.no_grad():
        p40 = model.predict_one_step(torch.zeros(3, 40), torch.from_numpy(states40), torch.from_numpy(sd40),
                                     torch.from_numpy(dens40), nums)
    G["C/s_cur"], G["C/s_delta"], G["C/dens"], G["C/nums"], G["C/s_pred"] = \
        states40, sd40, dens40, nums.numpy(), p40.numpy()
    G["C/rel"] = coo(ref_adjacency(model, torch.from_numpy(states40), torch.from_numpy(sd40), nums))

    # ---- case D: rollout + reward + action gradient, n_batch=2 x n_sample=3, N=60, T=4 ------------
    n_batch, n_sample, N, T = 2, 3, 60, 4
    st, dn = synthetic.make_pile_batch(n_batch, N, seed=3)
    acts = synthetic.random_actions(n_sample * n_batch, T, seed=3)
    goal = synthetic.make_goal("bar")
    goal_t = torch.from_numpy(goal)
    coords = torch.flip((goal_t < 0.5).nonzero(), dims=(1,)).float().numpy()
    goal_coor, _ = ref["utils"].fps_np(coords, min(5 * N, coords.shape[0]), 0)
    planner.particle_num = N
    act_t = torch.tensor(acts, requires_grad=True)
    with ref_harness.cpu_cuda_shims():
        out = planner.ptcl_model_rollout(torch.from_numpy(st), torch.from_numpy(dn), torch.zeros(n_batch, N),
                                         model, act_t)
    pred = out["model_rollout"]["state_pred"]                       # [Bt,T,N,3]
    obs = pred.reshape(n_sample * n_batch, 1, T, N, 3).permute(0, 2, 1, 3, 4)
    reward, next_r = planner.ptcl_evaluate_traj(obs, goal_t, torch.from_numpy(goal_coor).float())
    loss = torch.sum(-reward)
    loss.backward()
    G["D/s0"], G["D/dens"], G["D/acts"], G["D/goal_coor"] = st, dn, acts, goal_coor.astype(np.float32)
    G["D/goal_kind"] = np.array("bar")
    G["D/state_pred"] = pred.detach().numpy()
    G["D/reward"] = reward.detach().numpy()
    G["D/next_r"] = next_r.detach().numpy()
    G["D/act_grad"] = act_t.grad.numpy()
    # relation sets at every step, recomputed by the reference from the states it produced
    rels = []
    s = torch.from_numpy(st).repeat(n_sample, 1, 1)
    for t in range(T):
        with torch.no_grad():
            sdel = planner.gen_s_delta(s, act_t.detach()[:, t])
            rels.append(coo(ref_adjacency(model, s, sdel)))
            s = pred.detach()[:, t]
    for t, r in enumerate(rels):
        G["D/rel%d" % t] = r

    # ---- case E: MPPI pieces (numpy, global RNG) -----------------------------------------------
    np.random.seed(11)
    init = np.random.uniform(-4, 4, size=(5, 1, 4))
    np.random.seed(12)
    sampled = planner.sample_action_sequences(init, np.zeros(5), 16, None, None)
    rew = np.random.RandomState(13).uniform(-30, -5, size=(16, 1))
    G["E/init"], G["E/sampled"], G["E/reward"] = init, sampled, rew
    G["E/optimized"] = planner.optimize_action(sampled, rew)

    path = os.path.join(HERE, "golden_v1.npz")
    np.savez_compressed(path, **G)
    print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1e3), len(G), "arrays")


if __name__ == "__main__":
    main()
