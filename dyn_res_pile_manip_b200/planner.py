"""Drop-in replacement of the reference planner's particle path (planners.py:64-871).

`PlannerGD(config, env)` keeps the reference entry points -- `gen_s_delta`,
`ptcl_model_rollout`, `ptcl_evaluate_traj`, `sample_action_sequences`, `optimize_action`,
`trajectory_optimization_ptcl_multi_traj` -- with the same argument meaning, asserts and
returned dict, and adds `trajectory_optimization_mppi` (the MPPI composition the reference
ships in pieces but never wires up).  All tensor work runs in
libpilegnn; torch provides memory, streams, autograd glue and the process group.
"""
import time
from collections import OrderedDict

import numpy as np
import torch

from . import _lib, ops
from .propnet import PropNetDiffDenModel
from .rewards import GoalCache, config_reward_ptcl, goal_content_hash
from .synthetic import fps_np

DEBUG = False


def particle_num_to_iter_time(particle_num):
    """The authors' fitted ms-per-iteration model that caps the iteration count (planners.py:25-28)."""
    t = (2969.3971 - 69.923244 * particle_num + 1.8509846 * particle_num ** 2) / 200.
    return max(int(t), 1)


def general_rollout(planner, model_dy, s0, dens, attr, act_seqs):
    """T-step rollout for a model of any width (nf_effect != 64): the reference's own loop (planners.py:341-359) --
    pusher model, model step -- on the differentiable ops of the general-width engine; gradients reach the action
    sequences through torch's autograd over those two ops.  No weight gradients (planners.py:674 optimises actions only)."""
    from .propnet import general_step
    states = []
    s = s0
    for t in range(act_seqs.shape[1]):
        s_delta = ops.gen_s_delta(s, act_seqs[:, t], planner.pusher)
        s = general_step(s, s_delta, attr, dens, model_dy.model, None, None)
        states.append(s)
    return torch.stack(states, dim=1)


class _RolloutFn(torch.autograd.Function):
    """T-step rollout; differentiable w.r.t. the action sequences only (planners.py:674)."""

    @staticmethod
    def forward(ctx, act_seqs, planner, model_dy, s0, dens, attr):
        dev = s0.device
        acts = ops._f32(act_seqs.detach(), dev)
        Bt, T, _ = acts.shape
        N = s0.shape[1]
        need_grad = act_seqs.requires_grad
        net = model_dy.model
        wpack = net.packed_weights(dev)
        scratch = net.workspace.scratch(Bt, N, dev)
        tape = ops.new_tape(Bt, N, T, dev) if need_grad else None
        states = ops.rollout_forward_raw(wpack, attr, dens, s0, acts, planner.pusher, model_dy.adj_thresh, scratch, tape)
        # tensors go through save_for_backward (the output too: keeping it on ctx would be an output -> grad_fn ->
        # ctx reference cycle that only the garbage collector frees, with the tape hanging off it)
        if need_grad:
            ctx.save_for_backward(wpack, dens, s0, acts, tape, states)
        ctx.has_tape = need_grad
        ctx.planner, ctx.net = planner, net
        return states

    @staticmethod
    def backward(ctx, g_states):
        if not ctx.has_tape:
            raise RuntimeError("rollout was recorded without a tape (actions did not require grad)")
        wpack, dens, s0, acts, tape, states = ctx.saved_tensors
        Bt, T, N, _ = states.shape
        g = ops._f32(g_states).clone()          # consumed by the backward sweep
        g_act = ops.rollout_backward_raw(wpack, dens, s0, acts, ctx.planner.pusher, tape, states, g,
                                         ctx.net.workspace.bwd(Bt, N, states.device))
        return g_act, None, None, None, None, None


class _GDLoop:
    """Static device state of the GD planner's optimisation loop for one problem size.

    Every buffer an iteration touches is allocated once, so the two halves of an iteration --
    (rollout + last-step reward + best tracking) and (reward/rollout backward + Adam/clamp + counter) -- are
    fixed launch sequences without a host-side value in them: they are captured as two CUDA graphs and replayed
    n_iter times (two graphs so that the reference's rollout_time / optim_time split survives as event pairs
    around the replays)."""
    REW_CAP = 1024

    def __init__(self, key, net, device):
        rows, n_batch, N, T, M, Hh, Ww, n_batch1 = key[:8]
        self.key = key
        self.rows, self.n_batch, self.N, self.T, self.M = rows, n_batch, N, T, M
        self.n_batch1 = n_batch1                 # state variants per scene (n_batch = scenes * n_batch1)
        self.calls = 0
        f = dict(dtype=torch.float32, device=device)
        i32 = dict(dtype=torch.int32, device=device)
        self.acts = torch.zeros(rows, T, 4, **f)
        self.s0 = torch.zeros(rows, N, 3, **f)
        self.dens = torch.ones(rows, **f)
        self.attr = torch.zeros(rows, N, **f)
        self.states = torch.zeros(rows, T, N, 3, **f)
        self.g_states = torch.zeros(rows, T, N, 3, **f)
        self.reward = torch.zeros(rows, **f)
        self.argmin = torch.zeros(rows, M, **i32)
        self.g_reward = torch.full((rows,), -1.0, **f)            # d(sum(-reward)) / d reward
        self.g_act = torch.zeros(rows, T, 4, **f)
        self.exp_avg = torch.zeros(rows, T, 4, **f)
        self.exp_avg_sq = torch.zeros(rows, T, 4, **f)
        self.goal_img = torch.zeros(Hh, Ww, **f)
        self.goal_coor = torch.zeros(M, 2, **f)
        self.max_reward = torch.zeros(n_batch, **f)
        self.max_idx = torch.zeros(n_batch, **i32)
        self.best_actions = torch.zeros(n_batch, T, 4, **f)
        self.rew_mean = torch.zeros(n_batch // n_batch1, self.REW_CAP, **f)
        self.rew_std = torch.zeros(n_batch // n_batch1, self.REW_CAP, **f)
        self.iter = torch.zeros(1, **i32)
        self.general = not net.planner_engines      # nf_effect != 64: rollout through general_rollout + autograd
        if not self.general:
            self.scratch = net.workspace.scratch(rows, N, device)
            self.bwd_scratch = net.workspace.bwd(rows, N, device)
            self.tape = ops.new_tape(rows, N, T, device)
        self._graph_states = None
        self.sig = None
        self.g_fwd = self.g_bwd = None

    def reset(self, s0, dens, attr, acts, goal_img, goal_coor):
        self.s0.copy_(s0); self.dens.copy_(dens); self.attr.copy_(attr); self.acts.copy_(acts)
        self.goal_img.copy_(goal_img); self.goal_coor.copy_(goal_coor)
        self.exp_avg.zero_(); self.exp_avg_sq.zero_(); self.iter.zero_()
        self.max_reward.fill_(-float('inf')); self.max_idx.zero_(); self.best_actions.zero_()

    def _enqueue_fwd(self, c):
        N, T = self.N, self.T
        last = self.states[:, T - 1]                              # view: last-step states, stride T*N*3
        if self.general:
            leaf = self.acts.detach().requires_grad_(True)
            with torch.enable_grad():
                self._graph_states = (general_rollout(c['planner'], c['model_dy'], self.s0, self.dens, self.attr, leaf), leaf)
            self.states.copy_(self._graph_states[0].detach())
        else:
            ops.rollout_forward_raw(c['wpack'], self.attr, self.dens, self.s0, self.acts, c['pusher'], c['adj_thresh'],
                                    self.scratch, self.tape, out=self.states)
        # the loss only looks at the last step (reward_seqs = next_r[:, -1], planners.py:438)
        ops.reward_raw(last, self.rows, T * N * 3, N, self.goal_img, self.goal_coor, c['cam'], c['offset'], True,
                       want_argmin=True, out=self.reward, arg=self.argmin)
        # per state variant: keep the best trajectory seen so far (planners.py:721-727) + rew_mean / rew_std
        ops.gd_track(self.reward, self.acts, self.rows // self.n_batch, self.n_batch, T, self.max_reward, self.max_idx,
                     self.best_actions, self.rew_mean, self.rew_std, self.iter, stat_every=self.n_batch1,
                     stat_stride=self.REW_CAP)

    def _enqueue_bwd(self, c):
        N, T = self.N, self.T
        last = self.states[:, T - 1]
        # loss = sum(-reward): only the last step has an upstream gradient; earlier slices accumulate the
        # state gradients of the sweep and must start at zero
        if T > 1:
            self.g_states[:, :T - 1].zero_()
        ops.reward_backward_raw(last, self.rows, T * N * 3, N, self.goal_img, self.goal_coor, c['cam'], c['offset'],
                                True, self.g_reward, self.argmin, self.g_states[:, T - 1], T * N * 3, False)
        if self.general:
            states, leaf = self._graph_states
            self._graph_states = None
            (g_leaf,) = torch.autograd.grad([states], [leaf], [self.g_states])
            self.g_act.copy_(g_leaf)
        else:
            ops.rollout_backward_raw(c['wpack'], self.dens, self.s0, self.acts, c['pusher'], self.tape, self.states,
                                     self.g_states, self.bwd_scratch, out=self.g_act)
        ops.adam_clamp_dev(self.acts, self.g_act, self.exp_avg, self.exp_avg_sq, self.iter, c['lr'], c['lo'], c['hi'])
        ops.counter_add(self.iter, 1)

    def ensure_captured(self, sig, c):
        """(Re)capture when anything baked into the launches changed (weights buffer, engine, pusher, camera, box)."""
        if self.sig == sig and self.g_fwd is not None:
            return False
        self._enqueue_fwd(c)                    # eager warm-up (one-time kernel attribute setup)
        self._enqueue_bwd(c)
        torch.cuda.synchronize()
        g1, g2 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        with torch.cuda.graph(g1):
            self._enqueue_fwd(c)
        with torch.cuda.graph(g2):
            self._enqueue_bwd(c)
        self.g_fwd, self.g_bwd, self.sig = g1, g2, sig
        return True


class Planner(object):
    """reference planners.py:30-62 (attributes read from the environment)."""

    def __init__(self, config, env):
        self.config = config
        self.action_dim = 4
        self.global_scale = config['dataset']['global_scale']
        self.img_ch = 1
        self.n_his = config['train']['n_history']
        self.env = env
        self.cam_params = self.env.get_cam_params()
        self.is_real = self.env.is_real
        if not self.is_real:
            self.cam_extrinsic = self.env.get_cam_extrinsics()
        self.screenHeight = self.env.screenHeight
        self.screenWidth = self.env.screenWidth

    def trajectory_optimization(self, state_cur, obs_goal, model_dy, act_seq, n_sample, n_look_ahead, n_update_iter,
                                action_lower_lim, action_upper_lim, use_gpu):
        pass


class PlannerGD(Planner):

    def __init__(self, config, env):
        super(PlannerGD, self).__init__(config, env)
        if self.is_real:
            # real robot: the pusher model is gen_s_delta_irl (planners.py:259-300, dispatch :345-348)
            self.pusher = ops.Pusher.real(self.env.s2r_scale, self.env.wkspc_center_x, self.env.wkspc_center_y)
        else:
            self.cam12 = ops.cam_matrix12(self.cam_extrinsic)
            self.pusher = ops.Pusher.sim(self.cam_extrinsic, self.global_scale)
        self.goals = GoalCache()
        self.dist_group = None      # torch.distributed process group for sample-sharded planning
        self.device = torch.device('cuda')
        self._goal_coor_cache = {}
        self._gd_loops = OrderedDict()      # captured optimisation loops, most recently used last
        self._mppi_engines = OrderedDict()  # captured MPPI evaluations (engine.RolloutEngine), most recently used last
        self.use_graph = True               # False: launch every iteration's kernels one by one (debugging)
        self.capture_first_call = False     # True: capture the loop already when a problem size is seen the first time

    def reward_offset(self):
        """Pixel offset of the reward projection (planners.py:409-412): the real camera image is cropped."""
        if self.env.is_real:
            return (float(-self.env.crop_w_lower + self.env.crop_w_off), float(-self.env.crop_h_lower + self.env.crop_h_off))
        return (0., 0.)

    # ---- workspace box (planners.py:150-155, 756-760) ---------------------------------------------
    def action_box(self, cvx_l=0):
        r = self.env.cvx_region
        x_diff, y_diff = r[cvx_l, 1] - r[cvx_l, 0], r[cvx_l, 3] - r[cvx_l, 2]
        lo = np.array([r[cvx_l, 0], r[cvx_l, 2], r[cvx_l, 0] + x_diff * 0.15, r[cvx_l, 2] + y_diff * 0.15])
        hi = np.array([r[cvx_l, 1], r[cvx_l, 3], r[cvx_l, 1] - x_diff * 0.15, r[cvx_l, 3] - y_diff * 0.15])
        return lo, hi

    # ---- MPPI pieces (host numpy like the reference, planners.py:69-190) --------------------------
    def sample_action_sequences(self, init_act_seq, init_act_label_seq, n_sample, action_lower_lim, action_upper_lim,
                                noise_type="normal"):
        beta = self.config['mpc']['mppi']['beta_filter']
        nd = init_act_seq.ndim
        assert nd in (2, 3)
        act_seqs = np.stack([init_act_seq] * n_sample)
        resid = np.zeros((n_sample,) + init_act_seq.shape[1:])
        for i in range(self.n_his - 1, init_act_seq.shape[0]):
            if noise_type == "normal":
                sigma = self.config['mpc']['sigma'] * self.global_scale / 12.0
                noise = np.random.normal(0, sigma, resid.shape)
            elif noise_type == "uniform":
                sigma = 2.0 * self.global_scale / 12.0
                noise = np.random.uniform(-sigma, sigma, resid.shape)
            elif noise_type == "total_rand":
                noise = np.zeros(resid.shape)
            else:
                raise ValueError("unknown noise type: %s" % (noise_type))
            resid = beta * noise + resid * (1. - beta)
            act_seqs[:, i] += resid
            if nd == 2:
                lo, hi = self.action_box(int(init_act_label_seq[i]))
                act_seqs[:, i] = np.clip(act_seqs[:, i], lo, hi)
            else:
                lo, hi = self.action_box(0)       # only trajectory 0 is clipped (planners.py:161-167)
                act_seqs[:, i, 0] = np.clip(act_seqs[:, i, 0], lo, hi)
            if noise_type == 'total_rand':
                lo, hi = self.action_box(0)
                act_seqs[:, i, 0] = np.random.uniform(lo, hi, (n_sample, self.action_dim))
        return act_seqs

    def optimize_action(self, act_seqs, reward_seqs):
        """softmax(reward_weight * reward)-weighted mean of the sampled sequences (planners.py:549-561)."""
        w = self.config['mpc']['mppi']['reward_weight']
        assert len(act_seqs.shape) == 4
        n_sample, n_look_ahead, cvx_num, action_dim = act_seqs.shape
        dev = torch.device('cuda')
        out = np.zeros((n_look_ahead, cvx_num, action_dim))
        for i in range(cvx_num):
            a = torch.as_tensor(np.ascontiguousarray(act_seqs[:, :, i, :]), dtype=torch.float32, device=dev)
            r = torch.as_tensor(np.ascontiguousarray(reward_seqs[:, i]), dtype=torch.float32, device=dev)
            rec = ops.mppi_partials(r, a, w)
            out[:, i, :] = (rec[2:] / rec[1]).reshape(n_look_ahead, action_dim).cpu().numpy()
        return out

    # ---- pusher model (planners.py:192-257) --------------------------------------------------------
    def world2cam(self, world_pts):
        assert type(world_pts) == torch.Tensor
        m = torch.tensor(self.cam12, device=world_pts.device, dtype=world_pts.dtype).view(3, 4)
        return (world_pts @ m[:, :3].T + m[:, 3]) / self.global_scale

    def gen_s_delta(self, s_cur: torch.Tensor, action: torch.Tensor):
        assert type(s_cur) == torch.Tensor
        assert s_cur.shape[1:] == (self.particle_num, 3)
        assert s_cur.shape[0] == action.shape[0]
        assert type(action) == torch.Tensor
        if self.is_real:
            raise _lib.PileLibraryError("gen_s_delta needs the simulator camera; this planner was built for env.is_real")
        return ops.gen_s_delta(s_cur, action, self.pusher)

    def gen_s_delta_irl(self, s_cur: torch.Tensor, action: torch.Tensor):
        """Real-robot pusher model (planners.py:259-300): reads env.wkspc_center_x/y and env.s2r_scale."""
        assert type(s_cur) == torch.Tensor
        assert s_cur.shape[1:] == (self.particle_num, 3)
        assert s_cur.shape[0] == action.shape[0]
        assert type(action) == torch.Tensor
        pusher = self.pusher if self.is_real else ops.Pusher.real(self.env.s2r_scale, self.env.wkspc_center_x,
                                                                  self.env.wkspc_center_y)
        return ops.gen_s_delta(s_cur, action, pusher)

    # ---- rollout (planners.py:302-370) ---------------------------------------------------------------
    def ptcl_model_rollout(self, s_cur_tensor, s_param_tensor, a_cur_tensor, model_dy, act_seqs, enable_grad=True):
        n_sample_times_n_batch, T, action_dim = act_seqs.size()
        n_batch = s_cur_tensor.shape[0]
        n_sample = n_sample_times_n_batch // n_batch
        assert type(s_cur_tensor) == torch.Tensor
        assert type(a_cur_tensor) == torch.Tensor
        assert s_cur_tensor.shape[1] == self.particle_num
        assert s_cur_tensor.shape[2] == 3
        assert a_cur_tensor.shape[1] == self.particle_num
        assert type(act_seqs) == torch.Tensor
        if type(model_dy) != PropNetDiffDenModel:
            raise NotImplementedError
        ops._require_cuda(s_cur_tensor, "s_cur_tensor")
        dev = s_cur_tensor.device
        # flat row = sample * n_batch + b (state tiled n_sample times, planners.py:336-339)
        s0 = ops._f32(s_cur_tensor.detach()).repeat(n_sample, 1, 1)
        dens = ops._f32(s_param_tensor.detach(), dev).repeat(n_sample)
        attr = ops._f32(a_cur_tensor.detach(), dev).repeat(n_sample, 1)
        start = torch.cuda.Event(enable_timing=True)
        end = torch.cuda.Event(enable_timing=True)
        start.record()
        if model_dy.model.planner_engines:
            states = _RolloutFn.apply(act_seqs, self, model_dy, s0, dens, attr)
        else:
            states = general_rollout(self, model_dy, s0, dens, attr, ops._f32(act_seqs, dev))
        end.record()
        end.synchronize()          # the reference synchronises after every step (planners.py:357); here once
        return {'model_rollout': {'state_pred': states}, 'rollout_time': start.elapsed_time(end)}

    # ---- reward (planners.py:372-452) -------------------------------------------------------------------
    def ptcl_evaluate_traj(self, obs_seqs, obs_goal, obs_goal_coor_tensor, debug=False, funnel_dist=None,
                           distractor_df_fn=None, act_seqs_tensor=None, normalize_rew=True):
        assert type(obs_seqs) == torch.Tensor
        assert len(obs_seqs.shape) == 5
        assert obs_seqs.shape[3] == self.particle_num
        assert obs_seqs.shape[4] == 3
        assert type(obs_goal) == torch.Tensor
        assert len(obs_goal.shape) == 2
        assert obs_goal.shape[0] == self.screenHeight
        assert obs_goal.shape[1] == self.screenWidth
        if distractor_df_fn is not None:
            raise NotImplementedError("distractor reward is inactive in the reference MPC loop (env/flex_env.py:1048-1065)")
        n_sample, n_look_ahead, cvx_num, _, _ = obs_seqs.shape
        obs_future = obs_seqs.reshape(n_sample * n_look_ahead * cvx_num, self.particle_num, 3)
        next_r = config_reward_ptcl(obs_future, obs_goal, cam_params=self.cam_params, goal_coor=obs_goal_coor_tensor,
                                    normalize=normalize_rew, offset=self.reward_offset(), cache=self.goals)
        next_r = next_r.reshape(n_sample, n_look_ahead, cvx_num)
        reward_seqs = next_r[:, -1]
        assert reward_seqs.shape == (n_sample, cvx_num)
        return reward_seqs, next_r

    # ---- gradient-descent planner (planners.py:563-871) ---------------------------------------------
    def trajectory_optimization_ptcl_multi_traj(self, state_cur_np, state_param, attr_cur_np, obs_goal, model_dy, act_seq,
                                                act_label_seq, n_sample, n_look_ahead, n_update_iter, action_lower_lim,
                                                action_upper_lim, use_gpu=True, rollout_best_action_sequence=True,
                                                reward_params=None, funnel_dist=None, distractor_df_fn=None, gd_loop=1,
                                                time_lim=float('inf')):
        assert type(state_cur_np) == np.ndarray
        assert len(state_cur_np.shape) == 3
        assert state_cur_np.shape[0] == state_param.shape[0]
        assert state_cur_np.shape[2] == 3
        assert type(state_param) == np.ndarray
        return self._plan_scenes([state_cur_np], [state_param], [attr_cur_np], obs_goal, model_dy, act_seq, act_label_seq,
                                 n_sample, n_look_ahead, n_update_iter, use_gpu, rollout_best_action_sequence, gd_loop,
                                 time_lim)[0]

    def trajectory_optimization_ptcl_multi_scene(self, state_cur_list, state_param_list, attr_cur_list, obs_goal, model_dy,
                                                 act_seq, act_label_seq, n_sample, n_look_ahead, n_update_iter,
                                                 action_lower_lim=None, action_upper_lim=None, use_gpu=True,
                                                 rollout_best_action_sequence=True, gd_loop=1, time_lim=float('inf')):
        """Several independent planner calls with the same goal, initial action set, particle count and number of
        state variants in ONE optimisation loop -- the heaviest caller of the hot path, the Bayesian-optimisation data
        generation, evaluates every candidate resolution with five MPC episodes from the same start
        (data_gen/res_rgr_data.py:128-221, test_repeat = 5; :419-432).  Scene k's state variants simply become
        variants k*n_batch .. (k+1)*n_batch-1 of one batch (rows are independent through rollout, reward, backward
        and Adam), the per-variant best tracking is unchanged, and the vote / best-sequence re-rollout
        (planners.py:771-851) runs per scene.  Returns one result dict per scene, each bit-identical to what
        `trajectory_optimization_ptcl_multi_traj` returns for that scene alone."""
        assert len(state_cur_list) == len(state_param_list) == len(attr_cur_list) and len(state_cur_list) >= 1
        for st, sp in zip(state_cur_list, state_param_list):
            assert type(st) == np.ndarray and st.shape == state_cur_list[0].shape and st.shape[0] == sp.shape[0]
        return self._plan_scenes(state_cur_list, state_param_list, attr_cur_list, obs_goal, model_dy, act_seq, act_label_seq,
                                 n_sample, n_look_ahead, n_update_iter, use_gpu, rollout_best_action_sequence, gd_loop,
                                 time_lim)

    def _plan_scenes(self, state_cur_list, state_param_list, attr_cur_list, obs_goal, model_dy, act_seq, act_label_seq,
                     n_sample, n_look_ahead, n_update_iter, use_gpu, rollout_best_action_sequence, gd_loop, time_lim):
        time_lim = time_lim / 1000.0
        assert type(obs_goal) == np.ndarray
        assert len(obs_goal.shape) == 2
        assert type(act_seq) == np.ndarray
        assert len(act_seq.shape) == 3
        assert act_seq.shape[0] == act_label_seq.shape[0]
        assert len(act_label_seq.shape) == 1
        if not use_gpu:
            raise _lib.PileLibraryError("use_gpu=False: this planner has no CPU path")

        n_scene = len(state_cur_list)
        state_cur_np = np.concatenate(state_cur_list, axis=0)
        self.particle_num = state_cur_np.shape[1]
        n_batch1 = state_cur_list[0].shape[0]               # state variants per scene
        n_batch = state_cur_np.shape[0]                     # ... of the whole batch (scene-major)
        device = torch.device('cuda')
        state_cur_tensor = torch.tensor(state_cur_np, device=device, dtype=torch.float)
        attr_cur_tensor = torch.tensor(np.concatenate(attr_cur_list, axis=0), device=device, dtype=torch.float)
        obs_goal_tensor = torch.tensor(obs_goal, device=device, dtype=torch.float)
        goal_hash = goal_content_hash(obs_goal)          # one pass over the image for both goal caches
        obs_goal_coor_tensor = self.goal_coordinates(obs_goal, device, goal_hash)
        state_param_tensor = torch.from_numpy(np.concatenate(state_param_list, axis=0)).to(device=device, dtype=torch.float)

        n_act = act_seq.shape[0]
        traj_num = int(act_seq.shape[1])
        assert n_act == n_look_ahead
        assert traj_num == n_sample

        n_iter = min(n_update_iter, int(time_lim * 1000.0 / particle_num_to_iter_time(self.particle_num))) \
            if time_lim != float('inf') else n_update_iter
        n_iter = min(n_iter, _GDLoop.REW_CAP)

        # ---- static device state of the optimisation: no autograd graph, no host-side value and no host sync
        # inside an iteration, so each half of it is one CUDA-graph replay ------------------------------------
        N, T = self.particle_num, n_act
        rows = traj_num * n_batch                       # flat row = traj * n_batch + b  (planners.py:661-663)
        act_seqs = np.repeat(act_seq.transpose(1, 0, 2)[:, :, np.newaxis, :], n_batch, axis=0)
        act_seqs_tensor = torch.tensor(act_seqs, device=device, dtype=torch.float)     # [rows, T, 1, 4]
        net = model_dy.model
        lo, hi = self.action_box(0)
        ctx = {'wpack': net.packed_weights(device), 'pusher': self.pusher, 'adj_thresh': model_dy.adj_thresh,
               'cam': [float(v) for v in self.cam_params], 'offset': self.reward_offset(),
               'lr': self.config['mpc']['gd']['lr'], 'lo': [float(v) for v in lo], 'hi': [float(v) for v in hi],
               'planner': self, 'model_dy': model_dy}
        sig = (ctx['wpack'].data_ptr(), _lib.load().pile_get_tensor_cores(), self.pusher.signature(),
               float(model_dy.adj_thresh), tuple(ctx['cam']), ctx['offset'], float(ctx['lr']), tuple(ctx['lo']),
               tuple(ctx['hi']))
        start = time.time()
        timed = []
        done = 0
        loop = None
        try:
            M = int(obs_goal_coor_tensor.shape[0])
            key = (rows, n_batch, N, T, M, int(obs_goal_tensor.shape[0]), int(obs_goal_tensor.shape[1]), n_batch1, str(device))
            loop = self._gd_loops.pop(key, None)
            if loop is None:
                while len(self._gd_loops) >= 2:          # each loop owns a tape: keep the two most recent sizes
                    self._gd_loops.popitem(last=False)
                loop = _GDLoop(key, net, device)
            self._gd_loops[key] = loop
            goal_img = self.goals.shaped_np(obs_goal, obs_goal_tensor, goal_hash)
            inputs = (state_cur_tensor.repeat(traj_num, 1, 1), state_param_tensor.repeat(traj_num),
                      attr_cur_tensor.repeat(traj_num, 1), act_seqs_tensor.view(rows, T, 4), goal_img,
                      obs_goal_coor_tensor)
            loop.reset(*inputs)
            # a problem size seen for the first time runs launch by launch (the loop is GPU-bound, the host keeps up);
            # from the second call on the two captured graphs are replayed
            loop.calls += 1
            graph = (self.use_graph and n_iter > 0 and (loop.calls >= 2 or self.capture_first_call)
                     and not loop.general)          # the general-width rollout runs under autograd: no capture
            if graph and loop.ensure_captured(sig, ctx):
                loop.reset(*inputs)                      # the capture's warm-up iteration moved the actions
            for i in range(n_iter):
                e0, e1, e2 = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
                e0.record()
                if graph:
                    loop.g_fwd.replay()
                else:
                    loop._enqueue_fwd(ctx)
                e1.record()
                if graph:
                    loop.g_bwd.replay()
                else:
                    loop._enqueue_bwd(ctx)
                e2.record()
                timed.append((e0, e1, e2))
                done = i + 1
        except torch.cuda.OutOfMemoryError:
            # the reference's loop prints this and keeps what it has (planners.py:694-696, 748-750); every other
            # failure (bad sizes, launch errors) propagates as PileLibraryError
            print('OOM error')
        torch.cuda.synchronize()
        i = done - 1
        rollout_time = float(sum(a.elapsed_time(b) for a, b, _ in timed))
        optim_time = float(sum(b.elapsed_time(c) for _, b, c in timed))
        if loop is not None and done > 0:
            rew_mean_all = loop.rew_mean[:n_scene, :done].cpu().numpy()
            rew_std_all = loop.rew_std[:n_scene, :done].cpu().numpy()
            reward_all = loop.reward.reshape(n_sample, n_batch)
            act_seqs_tensor = loop.acts.view(rows, T, 1, 4)
            max_reward_all, max_idx_all = loop.max_reward.clone(), loop.max_idx.long()
            best_actions_all = loop.best_actions.clone()
        else:
            rew_mean_all = rew_std_all = np.zeros((n_scene, 0), dtype=np.float32)
            reward_all = torch.ones((n_sample, n_batch), device=device, dtype=torch.float)
            max_reward_all = -float('inf') * torch.ones(n_batch, device=device, dtype=torch.float)
            max_idx_all = torch.zeros(n_batch, device=device, dtype=torch.long)
            best_actions_all = torch.zeros((n_batch, n_act, self.action_dim), device=device, dtype=torch.float)
        reward_all_np = reward_all.data.cpu().numpy()
        act_all_np = act_seqs_tensor.data.cpu().numpy().reshape(traj_num, n_batch, T, 1, 4)

        results = []
        for k in range(n_scene):
            sl = slice(k * n_batch1, (k + 1) * n_batch1)
            rew_mean = np.zeros((1, n_update_iter * gd_loop), dtype=np.float32)
            rew_std = np.zeros((1, n_update_iter * gd_loop), dtype=np.float32)
            rew_mean[0, :done] = rew_mean_all[k]
            rew_std[0, :done] = rew_std_all[k]
            reward_seqs = reward_all_np[:, sl]
            act_seqs_k = act_all_np[:, sl].reshape(traj_num * n_batch1, T, 1, 4)
            # trajectories sharded over ranks (set planner.dist_group / traj_offset): the only exchange of the GD
            # planner is this one merge of the per-variant winners
            max_reward, max_reward_traj_idx, best_actions_of_samples = merge_best_across_ranks(
                max_reward_all[sl], max_idx_all[sl] + int(getattr(self, 'traj_offset', 0)), best_actions_all[sl],
                self.dist_group)
            # vote over state variants for the winning trajectory, then the best variant of it (planners.py:771-781)
            max_reward_traj_count = torch.bincount(max_reward_traj_idx)
            idx_best_act = torch.argmax(max_reward_traj_count).item()
            mr, mi = max_reward.cpu().numpy(), max_reward_traj_idx.cpu().numpy()
            idx_best_sample, reward_from_best_sample = -1, -float('inf')
            for j in range(n_batch1):
                if idx_best_act == mi[j] and mr[j] > reward_from_best_sample:
                    idx_best_sample, reward_from_best_sample = j, mr[j]
            act_seq_best = best_actions_of_samples.detach().cpu().numpy()[idx_best_sample][:, None, :]

            obs_seq_best = None
            reward_best = None
            next_r = None
            reward_best_idx = 0
            act_seq_out = act_seq_best.transpose(1, 0, 2)
            if rollout_best_action_sequence:
                assert act_seq_out.shape == (1, n_act, self.action_dim)
                act_seq_tensor = torch.from_numpy(act_seq_out).float().to(device)
                v0 = k * n_batch1
                out = self.ptcl_model_rollout(state_cur_tensor[v0:v0 + 1], state_param_tensor[v0:v0 + 1],
                                              attr_cur_tensor[v0:v0 + 1], model_dy, act_seq_tensor, enable_grad=True)
                obs_seq = out['model_rollout']['state_pred'].permute(1, 0, 2, 3).unsqueeze(0)
                reward_seq_best, next_seq_r = self.ptcl_evaluate_traj(obs_seq.contiguous(), obs_goal_tensor,
                                                                      obs_goal_coor_tensor)
                reward_best_idx = next_seq_r[:, 0].argmax()
                next_r = next_seq_r[reward_best_idx]
                reward_best = reward_seq_best[reward_best_idx]
                obs_seq_best = out['model_rollout']['state_pred'][reward_best_idx].detach().cpu().numpy()
            action_seq_future = act_seq_out[int(reward_best_idx)]
            total_time = time.time() - start
            results.append({'action_sequence': action_seq_future,
                            'action_full': act_seqs_k[:, 0, 0, :],
                            'reward_full': reward_seqs[:, 0],
                            'observation_sequence': obs_seq_best,
                            'observation_distractor_sequence': None,
                            'reward': None if reward_best is None else reward_best.detach().cpu().numpy(),
                            'next_r': None if next_r is None else next_r.detach().cpu().numpy(),
                            'rew_mean': rew_mean,
                            'rew_std': rew_std,
                            'times': {'total_time': total_time, 'rollout_time': rollout_time, 'optim_time': optim_time},
                            'iter_num': i})
        return results

    # ---- helpers -----------------------------------------------------------------------------------------
    def goal_coordinates(self, obs_goal, device, content_hash=None):
        """FPS-thinned (col,row) pixels of the goal region (planners.py:620-624).  On a CUDA device the
        sampling runs in libpilegnn (`pile_fps`, same picks as utils.fps_np); results are cached per goal."""
        g = np.asarray(obs_goal)
        key = (g.shape, self.particle_num, str(device), goal_content_hash(g) if content_hash is None else content_hash)
        hit = self._goal_coor_cache.get(key)
        if hit is not None:
            return hit
        rc = np.argwhere(g < 0.5)
        coords = rc[:, ::-1].astype(np.float32)
        count = min(self.particle_num * 5, coords.shape[0])
        if torch.device(device).type == 'cuda':
            picked, _, _ = ops.fps(torch.as_tensor(np.ascontiguousarray(coords), device=device), count, 0)
            out = picked.contiguous()
        else:
            picked, _ = fps_np(coords, count, 0)
            out = torch.tensor(picked, device=device, dtype=torch.float)
        if len(self._goal_coor_cache) > 8:
            self._goal_coor_cache.clear()
        self._goal_coor_cache[key] = out
        return out

    # ---- MPPI planner: sample -> rollout -> score -> softmax-weighted mean, sample-sharded ----------
    def trajectory_optimization_mppi(self, state_cur_np, state_param, attr_cur_np, obs_goal, model_dy, act_seq,
                                     n_sample, n_update_iter=1, seed=None):
        """act_seq [T,4] mean sequence -> dict(action_sequence [T,4], reward [n_sample local], ...).

        Composition of the reference's unwired MPPI pieces (planners.py:69-190, 549-561): every iteration
        draws n_sample filtered-noise perturbations of the mean sequence, rolls each out from state variant
        0, scores the last state, and replaces the mean by softmax(reward_weight*reward)-weighted samples.
        With a process group set (`self.dist_group`), each rank draws the SAME n_sample perturbations,
        evaluates its contiguous slice, and one all-gather of the (max, Z, A[T,4]) record per iteration
        merges the ranks -- the result is identical for any world size.
        """
        import torch.distributed as dist
        device = self.device
        self.particle_num = state_cur_np.shape[1]
        T = act_seq.shape[0]
        w = self.config['mpc']['mppi']['reward_weight']
        world, rank = 1, 0
        if self.dist_group is not None or (dist.is_available() and dist.is_initialized()):
            world, rank = dist.get_world_size(self.dist_group), dist.get_rank(self.dist_group)
        shard_size(n_sample, world)
        s0 = torch.tensor(state_cur_np[0:1], device=device, dtype=torch.float)
        dens = torch.tensor(np.asarray(state_param)[0:1], device=device, dtype=torch.float)
        attr = torch.tensor(attr_cur_np[0:1], device=device, dtype=torch.float)
        # one pass over the goal image for both goal caches; the image itself goes to the device only on a cache miss
        goal_hash = goal_content_hash(obs_goal)
        coor = self.goal_coordinates(obs_goal, device, goal_hash)
        goal_img = self.goals.shaped_np(obs_goal, None, goal_hash, device=device)
        mean = np.asarray(act_seq, dtype=np.float64).reshape(T, 1, 4)
        if seed is not None:
            np.random.seed(seed)
        rewards = None
        for _ in range(n_update_iter):
            sampled = self.sample_action_sequences(mean, np.zeros(T), n_sample, None, None)   # [n_sample,T,1,4]
            lo_i, hi_i = shard_bounds(n_sample, rank, world)
            mine = torch.tensor(sampled[lo_i:hi_i, :, 0, :], device=device, dtype=torch.float)
            with torch.no_grad():
                rewards, rec = self._mppi_evaluate(s0, dens, attr, model_dy, mine, goal_img, coor, w)
                if world > 1:
                    allrec = torch.empty(world * rec.numel(), device=device)
                    dist.all_gather_into_tensor(allrec, rec.contiguous(), group=self.dist_group)
                    rec = ops.mppi_combine(allrec.view(world, -1), T)
            mean = (rec[2:] / rec[1]).reshape(T, 1, 4).double().cpu().numpy()
        return {'action_sequence': mean[:, 0, :], 'reward': rewards.cpu().numpy(), 'record': rec.cpu().numpy()}

    def _mppi_evaluate(self, s0, dens, attr, model_dy, acts, goal_img, coor, reward_weight):
        """This rank's share of one MPPI iteration: roll `acts` [S,T,4] out from the single state variant, score the last
        state, reduce to the (max z, sum e^z, sum e^z act) record -> (rewards [S], record [2+4T]).  One captured CUDA
        graph per problem size (engine.RolloutEngine: T x 9 + 3 launches), reused across iterations and calls."""
        from .engine import RolloutEngine
        S, T = int(acts.shape[0]), int(acts.shape[1])
        N = int(s0.shape[1])
        key = (S, N, T, int(coor.shape[0]), tuple(goal_img.shape), float(reward_weight))
        eng = self._mppi_engines.pop(key, None)
        if eng is None:
            while len(self._mppi_engines) >= 2:
                self._mppi_engines.popitem(last=False)
            eng = RolloutEngine(model_dy, self, S, N, T, device=s0.device, reward_weight=reward_weight)
        self._mppi_engines[key] = eng
        eng.model_dy = model_dy
        # goal buffers and start state are re-loaded only when they are other objects than last time (the goal caches hand
        # out the same tensors while the goal is unchanged; s0 / dens / attr are built once per planner call)
        if getattr(eng, '_goal_src', None) is not goal_img or eng._coor_src is not coor:
            eng.set_goal_shaped(goal_img, coor)
            eng._goal_src, eng._coor_src = goal_img, coor
        if getattr(eng, '_state_src', None) is None or any(a is not b for a, b in zip(eng._state_src, (s0, dens, attr))):
            eng.load_state(s0, dens, attr)
            eng._state_src = (s0, dens, attr)
        eng.actions.copy_(acts)
        eng.evaluate()
        return eng.reward.clone(), eng.record.clone()


def merge_best_across_ranks(max_reward, traj_idx, best_actions, group=None):
    """Per state variant keep the best (reward, global trajectory index, action sequence) over all ranks.
    max_reward [n_batch], traj_idx [n_batch] (already global), best_actions [n_batch, T, 4]; ties go to the lower
    global trajectory index, which makes the result independent of the sharding.  No-op without a process group."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return max_reward, traj_idx, best_actions
    world = dist.get_world_size(group)
    n_batch = max_reward.shape[0]
    mine = torch.cat([max_reward.reshape(n_batch, 1).float(), traj_idx.reshape(n_batch, 1).float(),
                      best_actions.reshape(n_batch, -1).float()], dim=1).contiguous()
    everyone = torch.empty(world * mine.numel(), dtype=mine.dtype, device=mine.device)
    dist.all_gather_into_tensor(everyone, mine.reshape(-1), group=group)
    everyone = everyone.view(world, n_batch, -1)
    rew, idx = everyone[:, :, 0], everyone[:, :, 1]
    best_rew = rew.max(dim=0).values
    cand = torch.where(rew == best_rew[None], idx, torch.full_like(idx, float('inf')))
    win = cand.argmin(dim=0)                                   # rank holding the winner of each variant
    picked = everyone[win, torch.arange(n_batch, device=mine.device)]
    return picked[:, 0], picked[:, 1].long(), picked[:, 2:].reshape(best_actions.shape)


def shard_size(n_sample, world):
    if n_sample % world != 0:
        raise ValueError("n_sample=%d is not divisible by the %d ranks" % (n_sample, world))
    return n_sample // world


def shard_bounds(n_sample, rank, world):
    """Contiguous slice of the sample dimension owned by `rank` (SURVEY.md §8e partitioning)."""
    per = shard_size(n_sample, world)
    return rank * per, (rank + 1) * per
