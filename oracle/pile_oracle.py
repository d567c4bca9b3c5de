"""CPU oracle for the particle-GNN rollout hot path.  TEST INFRASTRUCTURE ONLY.

This file is the checker, never the product: only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s CPU-baseline / `--impl reference` legs may import it.  The shipped path
(`dyn_res_pile_manip_b200`) never imports anything from `oracle/` and has no CPU fallback.

It is a plain torch (fp32, CPU by default, device-agnostic so bench.py can also time the same dense
algorithm with torch's CUDA kernels as the "existing GPU path" comparator) restatement of the reference algorithm, written in the
reference's own dense formulation (one-hot receiver/sender matrices multiplied with
`bmm`) so that timing it is an honest stand-in for "the reference's PyTorch CPU path":

  relation build ............ model/gnn_dyn.py:221-251
  propagation network ....... model/gnn_dyn.py:147-198 (+ layer classes :6-111)
  pusher model .............. planners.py:192-257  (world2cam + gen_s_delta)
  horizon rollout ........... planners.py:302-370
  particle reward ........... env/flex_rewards.py:156-214, planners.py:372-452
  MPPI sampling / weighting . planners.py:69-190, 549-561

Parity pin: the reference has no tests or golden vectors of its own, so the
oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF: `tests/golden/make_golden.py`
imports /root/reference in the build container, runs it on seeded inputs and commits the
input/output vectors under `tests/golden/`; `tests/test_oracle_golden.py` checks this file
against them (relation sets bit-exact, floats to 1e-6), and `tests/test_oracle_vs_reference.py`
re-runs the live comparison whenever /root/reference is present.

Weights are passed as a plain dict keyed like the reference checkpoint
(`model.particle_encoder.model.0.weight`, ... see SURVEY.md §8b).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

PSTEP = 3          # model/gnn_dyn.py:160
MAX_REL = 10       # model/gnn_dyn.py:231
DENS_SCALE = 5000.  # model/gnn_dyn.py:158
PUSHER_W = 0.8 / 24.0   # planners.py:225
WIDTH_DECAY = 0.01      # planners.py:251


# --------------------------------------------------------------------------------------
# relations (model/gnn_dyn.py:221-251)
# --------------------------------------------------------------------------------------
def pair_sqdist(p):
    """dis[b,i,j] = sum_xyz (p[b,j]-p[b,i])^2, reduced in x,y,z order (gnn_dyn.py:224-230)."""
    diff = p[:, None, :, :] - p[:, :, None, :]          # sender j minus receiver i
    return (diff * diff).sum(-1)


def adjacency(s_cur, s_delta, adj_thresh, particle_nums=None):
    """bool [B,N,N]; adj[b,i,j] <=> j is one of the min(10,N) nearest of i AND within radius.

    Neighbours are searched on the *pushed* positions s_cur+s_delta (gnn_dyn.py:224-225).
    """
    p = s_cur + s_delta
    B, N, _ = p.shape
    dis = pair_sqdist(p)
    k = min(MAX_REL, N)
    nearest = torch.topk(dis, k=k, dim=2, largest=False).indices
    in_topk = torch.zeros_like(dis).scatter_(2, nearest, 1.0) > 0
    in_radius = (dis - adj_thresh * adj_thresh) < 0
    adj = in_topk & in_radius
    if particle_nums is not None:                       # gnn_dyn.py:238-241
        for b in range(B):
            n = int(particle_nums[b])
            adj[b, n:, :] = False
            adj[b, :, n:] = False
    return adj


def edge_lists(adj):
    """Per-sample (recv, send) int32 arrays in `nonzero()` order (row-major: recv, then send)."""
    out = []
    for b in range(adj.shape[0]):
        rs = adj[b].nonzero()
        out.append((rs[:, 0].to(torch.int32).cpu().numpy().copy(), rs[:, 1].to(torch.int32).cpu().numpy().copy()))
    return out


def one_hot_relations(adj, dtype=torch.float32):
    """Dense Rr, Rs [B, n_rel, N] exactly as the reference lays them out (gnn_dyn.py:242-251): relation slots of a
    sample in nonzero() order.  Built for the whole batch at once (the reference loops over samples in Python; the
    result is identical)."""
    B, N, _ = adj.shape
    counts = adj.sum(dim=(1, 2))
    n_rel = int(counts.max())
    Rr = torch.zeros(B, n_rel, N, dtype=dtype, device=adj.device)
    Rs = torch.zeros(B, n_rel, N, dtype=dtype, device=adj.device)
    brs = adj.nonzero()                                   # sorted by (b, recv, send)
    first = torch.cumsum(counts, 0) - counts              # index of each sample's first relation
    slot = torch.arange(brs.shape[0], device=adj.device) - first[brs[:, 0]]
    Rr[brs[:, 0], slot, brs[:, 1]] = 1
    Rs[brs[:, 0], slot, brs[:, 2]] = 1
    return Rr, Rs


# --------------------------------------------------------------------------------------
# propagation network forward (model/gnn_dyn.py:147-198)
# --------------------------------------------------------------------------------------
def _lin(W, prefix, x):
    return F.linear(x, W[prefix + ".weight"], W[prefix + ".bias"])


def propnet_forward(W, a_cur, s_cur, s_delta, Rr, Rs, particle_dens):
    """Dense one-hot formulation; returns s_pred [B,N,3]."""
    B, N = a_cur.shape
    n_rel = Rr.shape[1]
    d = particle_dens / DENS_SCALE
    d_node = d.view(B, 1, 1).expand(B, N, 1)
    d_rel = d.view(B, 1, 1).expand(B, n_rel, 1)
    attr = a_cur.unsqueeze(-1)

    # particle encoder: [s_delta(3), attr(1), dens(1)] -> H   (gnn_dyn.py:174-176, :43-48)
    x = torch.cat([s_delta, attr, d_node], dim=2)
    x = torch.relu(_lin(W, "model.particle_encoder.model.0", x))
    p_enc = torch.relu(_lin(W, "model.particle_encoder.model.2", x))

    # relation encoder: [attr_r, attr_s, s_r - s_s, dens] -> H   (gnn_dyn.py:166-171, 179-180)
    y = torch.cat([Rr.bmm(attr), Rs.bmm(attr), Rr.bmm(s_cur) - Rs.bmm(s_cur), d_rel], dim=2)
    y = torch.relu(_lin(W, "model.relation_encoder.model.0", y))
    y = torch.relu(_lin(W, "model.relation_encoder.model.2", y))
    r_enc = torch.relu(_lin(W, "model.relation_encoder.model.4", y))

    effect = p_enc
    RrT = Rr.transpose(1, 2)
    for _ in range(PSTEP):                               # gnn_dyn.py:182-193
        z = torch.cat([r_enc, Rr.bmm(effect), Rs.bmm(effect), d_rel], dim=2)
        e_rel = torch.relu(_lin(W, "model.relation_propagator.linear", z))
        agg = RrT.bmm(e_rel)
        z = torch.cat([p_enc, agg, d_node], dim=2)
        effect = torch.relu(_lin(W, "model.particle_propagator.linear", z) + effect)

    h = torch.relu(_lin(W, "model.particle_predictor.linear_0", effect))   # gnn_dyn.py:196-198
    return _lin(W, "model.particle_predictor.linear_1", h) + s_cur


def predict_one_step(W, adj_thresh, a_cur, s_cur, s_delta, particle_dens, particle_nums=None,
                     return_adj=False):
    """gnn_dyn.py:209-254."""
    adj = adjacency(s_cur.detach(), s_delta.detach(), adj_thresh, particle_nums)
    Rr, Rs = one_hot_relations(adj, s_cur.dtype)
    out = propnet_forward(W, a_cur, s_cur, s_delta, Rr, Rs, particle_dens)
    return (out, adj) if return_adj else out


def training_loss(W, adj_thresh, states, states_delta, attrs, particle_dens, particle_nums):
    """train/train_gnn_dyn.py:150-192: roll the model over states_delta [B, n_roll, N, 3] from states[:, 0], per
    sample MSE against states[:, t+1] over the first particle_nums[b] particles, mean over n_roll * B."""
    B, n_roll = states_delta.shape[0], states_delta.shape[1]
    s_cur, a_cur = states[:, 0], attrs[:, 0]
    loss = 0.
    for t in range(n_roll):
        s_pred = predict_one_step(W, adj_thresh, a_cur, s_cur, states_delta[:, t], particle_dens, particle_nums)
        for b in range(B):
            n = int(particle_nums[b])
            loss = loss + torch.mean((s_pred[b, :n] - states[b, t + 1, :n]) ** 2)
        s_cur = s_pred
    return loss / (n_roll * B)


# --------------------------------------------------------------------------------------
# pusher model (planners.py:192-257)
# --------------------------------------------------------------------------------------
def world_to_cam_matrix(cam_extrinsic):
    """The 4x4 the reference rebuilds on every call (planners.py:197-203), as float32."""
    flip = np.diag([1.0, -1.0, -1.0, 1.0])
    m = np.linalg.inv(np.matmul(np.linalg.inv(np.asarray(cam_extrinsic, dtype=np.float64)), flip))
    return torch.tensor(m).float()


def world2cam(cam_extrinsic, global_scale, pts):
    M = world_to_cam_matrix(cam_extrinsic).to(pts.device)
    homog = torch.cat([pts, torch.ones(pts.shape[0], 1, dtype=pts.dtype, device=pts.device)], dim=1)
    return torch.matmul(M, homog.T).T[:, :3] / global_scale


def gen_s_delta(cam_extrinsic, global_scale, s_cur, action):
    """s_cur [B,N,3], action [B,4]=(sx,sy,ex,ey) -> s_delta [B,N,3]   (planners.py:211-257)."""
    zero = torch.zeros(action.shape[0], 1, dtype=action.dtype, device=action.device)
    start = world2cam(cam_extrinsic, global_scale, torch.cat([action[:, 0:1], zero, -action[:, 1:2]], 1))
    end = world2cam(cam_extrinsic, global_scale, torch.cat([action[:, 2:3], zero, -action[:, 3:4]], 1))
    push = end - start
    length = torch.linalg.norm(push, dim=1)
    u = push / torch.linalg.norm(push, dim=1, keepdim=True)
    v = torch.cat([-u[:, 1:2], u[:, 0:1], torch.zeros_like(u[:, 0:1])], dim=1)
    rel = s_cur - start[:, None, :]
    across = (rel * v[:, None, :]).sum(-1)
    along = (rel * u[:, None, :]).sum(-1)
    hard = ((along < length[:, None]) & (along > 0.0)).float()
    excess = torch.maximum(torch.clamp(-PUSHER_W - across, min=0.), torch.clamp(across - PUSHER_W, min=0.))
    soft = torch.exp(-excess / WIDTH_DECAY)
    to_end = ((end[:, None, :] - s_cur) * u[:, None, :]).sum(-1)
    return to_end[..., None] * u[:, None, :] * hard[..., None] * soft[..., None]


IRL_PUSHER_W = 0.048     # planners.py:279
IRL_HEIGHT = 0.88        # planners.py:276


def gen_s_delta_irl(s_cur, action, wkspc_center_x, wkspc_center_y, s2r_scale):
    """Real-robot pusher parametrisation (planners.py:259-300, behind env.is_real): the push end points come from
    the action scaled by s2r_scale at a fixed height, the particles are shifted by the workspace centre.
    s_cur [B,N,3], action [B,4]=(sx,sy,ex,ey) -> s_delta [B,N,3]."""
    shift = torch.tensor([wkspc_center_x, wkspc_center_y, 0.0], dtype=s_cur.dtype, device=s_cur.device)
    p = s_cur - shift
    h = torch.full((action.shape[0], 1), IRL_HEIGHT, dtype=action.dtype, device=action.device)
    start = torch.cat([action[:, 0:1] / s2r_scale, -action[:, 1:2] / s2r_scale, h], dim=1)
    end = torch.cat([action[:, 2:3] / s2r_scale, -action[:, 3:4] / s2r_scale, h], dim=1)
    push = end - start
    length = torch.linalg.norm(push, dim=1)
    u = push / torch.linalg.norm(push, dim=1, keepdim=True)
    v = torch.cat([-u[:, 1:2], u[:, 0:1], torch.zeros_like(u[:, 0:1])], dim=1)
    rel = p - start[:, None, :]
    across = (rel * v[:, None, :]).sum(-1)
    along = (rel * u[:, None, :]).sum(-1)
    hard = ((along < length[:, None]) & (along > 0.0)).float()
    excess = torch.maximum(torch.clamp(-IRL_PUSHER_W - across, min=0.), torch.clamp(across - IRL_PUSHER_W, min=0.))
    soft = torch.exp(-excess / WIDTH_DECAY)
    to_end = ((end[:, None, :] - p) * u[:, None, :]).sum(-1)
    return to_end[..., None] * u[:, None, :] * hard[..., None] * soft[..., None]


# --------------------------------------------------------------------------------------
# rollout (planners.py:302-370)
# --------------------------------------------------------------------------------------
def rollout(W, adj_thresh, cam_extrinsic, global_scale, s0, dens, attr, act_seqs, return_adj=False):
    """s0 [n_batch,N,3], dens [n_batch], attr [n_batch,N], act_seqs [Bt,T,4] -> [Bt,T,N,3].

    Flat row index = sample * n_batch + b (state tiled n_sample x, planners.py:336-339).
    """
    Bt, T, _ = act_seqs.shape
    reps = Bt // s0.shape[0]
    s = s0.repeat(reps, 1, 1)
    dn = dens.repeat(reps)
    at = attr.repeat(reps, 1)
    states, adjs = [], []
    for t in range(T):
        sd = gen_s_delta(cam_extrinsic, global_scale, s, act_seqs[:, t])
        s, adj = predict_one_step(W, adj_thresh, at, s, sd, dn, return_adj=True)
        states.append(s)
        adjs.append(adj)
    out = torch.stack(states, dim=1)
    return (out, adjs) if return_adj else out


# --------------------------------------------------------------------------------------
# reward (env/flex_rewards.py:156-214)
# --------------------------------------------------------------------------------------
def shaped_goal_image(goal_np):
    """goal' = goal - DT(goal<0.5) - min   (flex_rewards.py:171-176); host side, cv2."""
    import cv2
    g = np.asarray(goal_np, dtype=np.float32)
    inside = cv2.distanceTransform((g < 0.5).astype(np.uint8), cv2.DIST_L2, 5)
    g = g - inside
    return g - g.min()


def reward_ptcl(state, goal, cam_params, goal_coor, normalize=True, offset=(0., 0.)):
    """state [B,N,3], goal [H,W] tensor, goal_coor [M,2]=(col,row) -> [B]."""
    B, N, _ = state.shape
    H, Wd = goal.shape
    fx, fy, cx, cy = cam_params
    img = torch.from_numpy(shaped_goal_image(goal.detach().cpu().numpy())).to(device=state.device, dtype=state.dtype)
    px = state[..., 0] * fx / state[..., 2] + cx + offset[0]
    py = state[..., 1] * fy / state[..., 2] + cy + offset[1]
    pix = torch.stack([px, py], dim=-1)                  # (col,row)
    grid = (pix / H * 2 - 1).unsqueeze(1)                # both axes over H (flex_rewards.py:197)
    looked = F.grid_sample(img.expand(B, 1, H, Wd), grid, padding_mode="border", align_corners=False)
    r = looked.reshape(B, N).sum(1)
    d = torch.norm(goal_coor[None, :, None, :] - pix[:, None, :, :], dim=3)   # [B,M,N]
    r = r + d.min(dim=2).values.sum(1)
    if normalize:
        r = r / N
    return -r


def evaluate_traj(obs_seqs, goal, cam_params, goal_coor, normalize=True):
    """obs_seqs [n,T,1,N,3] -> (reward [n,1] = last step, next_r [n,T,1])   (planners.py:372-452)."""
    n, T, C, N, _ = obs_seqs.shape
    r = reward_ptcl(obs_seqs.reshape(n * T * C, N, 3), goal, cam_params, goal_coor, normalize)
    r = r.reshape(n, T, C)
    return r[:, -1], r


# --------------------------------------------------------------------------------------
# MPPI pieces (planners.py:69-190, 549-561) -- numpy float64 like the reference
# --------------------------------------------------------------------------------------
def action_box(cvx_region, inset=0.15):
    """(lower[4], upper[4]); the push end point is inset 15 % (planners.py:150-155)."""
    l, r, b, t = [float(v) for v in cvx_region[0]]
    dx, dy = r - l, t - b
    return (np.array([l, b, l + dx * inset, b + dy * inset]),
            np.array([r, t, r - dx * inset, t - dy * inset]))


def sample_action_sequences(init_act_seq, n_sample, sigma, beta, cvx_region, n_his=1, rng=np.random):
    """init [T,traj,4] -> [n_sample,T,traj,4]; filtered Gaussian noise, only traj 0 clipped."""
    T, traj, A = init_act_seq.shape
    seqs = np.stack([init_act_seq] * n_sample).astype(np.float64)
    resid = np.zeros((n_sample, traj, A))
    lo, hi = action_box(cvx_region)
    for t in range(n_his - 1, T):
        noise = rng.normal(0, sigma, (n_sample, traj, A))
        resid = beta * noise + (1. - beta) * resid
        seqs[:, t] += resid
        seqs[:, t, 0] = np.clip(seqs[:, t, 0], lo, hi)
    return seqs


def mppi_optimize_action(act_seqs, reward_seqs, reward_weight):
    """act_seqs [n,T,C,4], reward [n,C] -> softmax-weighted mean [T,C,4]   (planners.py:549-561)."""
    n, T, C, A = act_seqs.shape
    out = np.zeros((T, C, A))
    for c in range(C):
        z = reward_weight * reward_seqs[:, c]
        w = np.exp(z - z.max())
        w = w / w.sum()
        out[:, c] = (w[:, None, None] * act_seqs[:, :, c]).sum(0)
    return out


def mppi_record(reward, acts, reward_weight):
    """(max z, sum exp(z-max), sum exp(z-max)*act) of one shard; acts [n,T,4] -> float64 [2+4T]."""
    z = reward_weight * np.asarray(reward, dtype=np.float64)
    m = z.max()
    w = np.exp(z - m)
    return np.concatenate([[m, w.sum()], (w[:, None, None] * np.asarray(acts, dtype=np.float64)).sum(0).reshape(-1)])


def mppi_merge(records):
    """log-sum-exp merge of shard records; merged[2:]/merged[1] equals mppi_optimize_action on the union."""
    records = np.asarray(records, dtype=np.float64)
    m = records[:, 0].max()
    sc = np.exp(records[:, 0] - m)
    return np.concatenate([[m, (sc * records[:, 1]).sum()], (sc[:, None] * records[:, 2:]).sum(0)])


# --------------------------------------------------------------------------------------
# helpers shared by tests / bench baseline
# --------------------------------------------------------------------------------------
def weights_from_seed(seed=0, nf=64):
    """Random-init weights with nn.Linear's default init, in the reference's construction order
    (particle_encoder, relation_encoder, particle_propagator, relation_propagator,
    particle_predictor: gnn_dyn.py:127-145) so `torch.manual_seed(seed)` reproduces the
    reference model's own initial state_dict."""
    g = torch.Generator().manual_seed(seed)
    torch_state = torch.get_rng_state()
    torch.manual_seed(seed)
    try:
        spec = [("model.particle_encoder.model.0", 5, nf), ("model.particle_encoder.model.2", nf, nf),
                ("model.relation_encoder.model.0", 6, nf), ("model.relation_encoder.model.2", nf, nf),
                ("model.relation_encoder.model.4", nf, nf),
                ("model.particle_propagator.linear", 2 * nf + 1, nf),
                ("model.relation_propagator.linear", 3 * nf + 1, nf),
                ("model.particle_predictor.linear_0", nf, nf), ("model.particle_predictor.linear_1", nf, 3)]
        W = {}
        for name, fan_in, fan_out in spec:
            lin = torch.nn.Linear(fan_in, fan_out)
            W[name + ".weight"] = lin.weight.detach().clone()
            W[name + ".bias"] = lin.bias.detach().clone()
    finally:
        torch.set_rng_state(torch_state)
    del g
    return W


def flops_per_sample_step(N, E, H=64):
    """(F_ref, F_alg) of SURVEY.md §8d."""
    f_ref = 2 * (N * (8 * H * H + 11 * H) + E * (11 * H * H + 9 * H))
    f_alg = 2 * (N * (14 * H * H + 11 * H) + E * (3 * H * H + 6 * H))
    return f_ref, f_alg
