"""The planner-side half of one closed-loop MPC step, without a simulator behind it.

`FlexEnv.step_subgoal_ptcl` (reference env/flex_env.py:953-1141) alternates simulator calls with planner-side
work: re-sample the observation into particles 30 times (:1028, :1086), take the particle density from the
covering radius (:1030, :1087), score the current state against the goal (:1031-1036, :1100-1104), optimise the
push sequence (:1048-1066) and shift the warm start (:1112-1115).  `MPCStep.plan` is that planner-side work for one
step, from the RGB-D observation to the action to execute, all on the GPU; stepping the simulator with the returned
action and rendering the next observation stays with the caller.
"""
import numpy as np
import torch

from . import observation
from .rewards import config_reward_ptcl


class MPCStep:
    def __init__(self, planner, model_dy, env, batch_size=30, res_rgr=None, resolution_buckets=None):
        """planner: PlannerGD; env: anything with get_cam_params() and global_scale (FlexEnv's interface);
        res_rgr: optional `regressor.MPCResRgrNoPool` -- the dynamic-resolution loop of step_subgoal_ptcl with
        auto_particle_r=True (flex_env.py:981-998, 1080-1090): when `plan` is called without a particle count the
        regressor picks it from (foreground mask, goal mask).  resolution_buckets: optional ascending list of
        allowed counts; the regressor's output is snapped to the nearest one (fewer distinct problem sizes means
        the planner's captured loops are reused more often).  Default: the reference's behaviour, the raw count."""
        self.planner, self.model_dy, self.env = planner, model_dy, env
        self.batch_size = int(batch_size)
        self.res_rgr = res_rgr
        self.resolution_buckets = None if resolution_buckets is None else sorted(int(b) for b in resolution_buckets)

    def select_resolution(self, obs, subgoal):
        """particle_num = res_rgr.infer_param(fg_mask, subgoal_mask)  (flex_env.py:993-997)."""
        if self.res_rgr is None:
            raise ValueError("MPCStep was built without a resolution regressor: pass particle_num")
        fg_mask = (np.asarray(obs)[..., -1] / self.env.global_scale < 0.599 / 0.8).astype(np.float32)
        subgoal_mask = (np.asarray(subgoal) < 0.5).astype(np.float32)
        n = int(self.res_rgr.infer_param(fg_mask, subgoal_mask))
        if self.resolution_buckets:
            n = min(self.resolution_buckets, key=lambda b: (abs(b - n), b))
        if n < 1:
            raise ValueError("resolution regressor returned %d particles" % n)
        return n

    def observe(self, obs, particle_num, init_idx=None, seed=None):
        """obs [H,W,5] -> (particles [batch,N,3] float64, particle_den [batch]) as flex_env.py:1028-1030."""
        obs_cur, particle_r = observation.obs2ptcl_fixed_num_batch(
            obs, particle_num, self.batch_size, self.env.get_cam_params(), self.env.global_scale, init_idx, seed)
        particle_den = np.array([1 / (particle_r * particle_r)])[0]
        return obs_cur, particle_den

    def reward(self, obs_cur, subgoal):
        """Normalised reward of the first re-sampling against the goal image (flex_env.py:1100-1104)."""
        particle_num = obs_cur.shape[1]
        goal = torch.from_numpy(np.asarray(subgoal)).float().cuda()
        self.planner.particle_num = particle_num
        coor = self.planner.goal_coordinates(np.asarray(subgoal), goal.device)
        state = torch.from_numpy(obs_cur).float().cuda().reshape(-1, particle_num, 3)
        return config_reward_ptcl(state, goal, cam_params=self.env.get_cam_params(), goal_coor=coor,
                                  normalize=True)[0].item()

    def plan(self, obs, subgoal, particle_num=None, action_seq_mpc_init=None, action_label_seq_mpc_init=None,
             n_sample=None, n_look_ahead=1, n_update_iter=100, action_lower_lim=None, action_upper_lim=None, gd_loop=1,
             time_lim=float('inf'), reward_params=None, init_idx=None, seed=None):
        """One MPC step: -> dict(action, traj_opt_out, obs_cur, particle_den, particle_num, reward,
        action_seq_mpc_init, action_label_seq_mpc_init) where the last two are the warm start of the next step.
        particle_num=None: the resolution regressor selects it (dynamic resolution)."""
        if particle_num is None:
            particle_num = self.select_resolution(obs, subgoal)
        if n_sample is None:
            n_sample = action_seq_mpc_init.shape[1]
        obs_cur, particle_den = self.observe(obs, particle_num, init_idx, seed)
        attr_cur = np.zeros((obs_cur.shape[0], particle_num))
        out = self.planner.trajectory_optimization_ptcl_multi_traj(
            obs_cur, particle_den, attr_cur, obs_goal=subgoal, model_dy=self.model_dy,
            act_seq=action_seq_mpc_init[:n_look_ahead],
            act_label_seq=action_label_seq_mpc_init[:n_look_ahead] if action_label_seq_mpc_init is not None else None,
            n_sample=n_sample, n_look_ahead=n_look_ahead, n_update_iter=n_update_iter,
            action_lower_lim=action_lower_lim, action_upper_lim=action_upper_lim, use_gpu=True,
            rollout_best_action_sequence=True, reward_params=reward_params, gd_loop=gd_loop, time_lim=time_lim)
        nxt = observation.shift_warm_start(action_seq_mpc_init, out['action_full'], n_look_ahead)
        nxt_label = action_label_seq_mpc_init
        if action_seq_mpc_init.shape[0] > 1 and action_label_seq_mpc_init is not None:
            nxt_label = action_label_seq_mpc_init[1:]
        return {'action': out['action_sequence'][0], 'traj_opt_out': out, 'obs_cur': obs_cur,
                'particle_den': particle_den, 'particle_num': int(particle_num), 'reward': self.reward(obs_cur, subgoal),
                'action_seq_mpc_init': nxt, 'action_label_seq_mpc_init': nxt_label}
