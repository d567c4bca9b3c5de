"""Static-shape evaluation engine: rollout over T + last-step reward + MPPI record as ONE CUDA graph.

The reference synchronises the host after every horizon step (planners.py:357) and builds
data-dependent shapes (`n_rel.item()`, gnn_dyn.py:243).  Here every buffer is sized by
(rows, N, T) up front (CSR with a fixed 10N edge capacity), so the T*9+3 kernel launches of one
planner evaluation (T*6+3 on the FP32 engine) are captured once with `torch.cuda.CUDAGraph` and replayed per iteration.
"""
import numpy as np
import torch

from . import _lib, ops
from .rewards import shape_goal_image


class RolloutEngine:
    """rows = n_sample * n_batch rollouts of horizon T over N particles on one GPU."""

    # kernels per model step: FP32 engine nbr_search, node_encode, edge_encode, 3 x propagate; tensor engine
    # nbr_search (also writes the relation-encoder input rows), node_encode_tc, edge_encode, 3 x (edge_agg +
    # node_update_tc)
    LAUNCHES_PER_MODEL_STEP = {0: 6, 1: 9, 2: 9}

    def __init__(self, model_dy, planner, rows, N, T, device=None, goal=None, goal_coor=None, use_graph=True,
                 reward_weight=None):
        self.model_dy, self.planner = model_dy, planner
        self.rows, self.N, self.T = int(rows), int(N), int(T)
        self.device = torch.device(device if device is not None else "cuda")
        # nf_effect != 64 rolls out step by step on the general-width engine (planner.general_rollout): no capture
        self.general = not model_dy.model.planner_engines
        self.use_graph = use_graph and not self.general
        dev = self.device
        self.actions = torch.zeros(rows, T, 4, dtype=torch.float32, device=dev)
        self.states = torch.empty(rows, T, N, 3, dtype=torch.float32, device=dev)
        self.s0 = torch.zeros(rows, N, 3, dtype=torch.float32, device=dev)
        self.dens = torch.ones(rows, dtype=torch.float32, device=dev)
        self.attr = torch.zeros(rows, N, dtype=torch.float32, device=dev)
        self.reward = torch.zeros(rows, dtype=torch.float32, device=dev)
        self.record = torch.zeros(2 + 4 * T, dtype=torch.float32, device=dev)
        self.reward_weight = planner.config['mpc']['mppi']['reward_weight'] if reward_weight is None else reward_weight
        lib = _lib.load()
        self._parts = torch.zeros(lib.pile_mppi_num_chunks(rows), 2 + 4 * T, dtype=torch.float32, device=dev)
        self.scratch = None if self.general else torch.empty(lib.pile_step_scratch_bytes(rows, N), dtype=torch.uint8, device=dev)
        self.goal_img = self.goal_coor = None
        if goal is not None:
            self.set_goal(goal, goal_coor)
        self._graph = None
        self._wpack = None
        self._mode = None

    # ---- inputs ---------------------------------------------------------------------------------
    def set_goal(self, goal, goal_coor=None):
        g = torch.as_tensor(goal, dtype=torch.float32, device=self.device)
        self.goal_img = shape_goal_image(g)
        if goal_coor is None:
            self.planner.particle_num = self.N
            goal_coor = self.planner.goal_coordinates(np.asarray(g.cpu()), self.device)
        self.goal_coor = ops._f32(goal_coor, self.device)
        self._graph = None

    def set_goal_shaped(self, goal_img, goal_coor):
        """Goal already shaped (rewards.shape_goal_image) and thinned on the device.  Same shapes as before: copied into
        the buffers the captured graph reads (the capture stays valid); otherwise the buffers are replaced."""
        if self.goal_img is not None and self.goal_img.shape == goal_img.shape and self.goal_coor.shape == goal_coor.shape:
            self.goal_img.copy_(goal_img)
            self.goal_coor.copy_(goal_coor)
            return
        self.goal_img = goal_img.detach().clone()
        self.goal_coor = ops._f32(goal_coor, self.device).clone()
        self._graph = None

    def load_state(self, s0, dens, attr=None):
        """s0 [n_batch,N,3], dens [n_batch] (+attr [n_batch,N]) tiled to rows: row = sample*n_batch + b."""
        s0 = ops._f32(s0, self.device)
        n_batch = s0.shape[0]
        assert self.rows % n_batch == 0 and s0.shape[1] == self.N
        reps = self.rows // n_batch
        self.s0.copy_(s0.repeat(reps, 1, 1))
        self.dens.copy_(ops._f32(dens, self.device).repeat(reps))
        if attr is not None:
            self.attr.copy_(ops._f32(attr, self.device).repeat(reps, 1))

    # ---- one evaluation --------------------------------------------------------------------------
    def launches_per_eval(self):
        n = self.T * self.LAUNCHES_PER_MODEL_STEP[_lib.load().pile_get_tensor_cores()]
        if self.goal_img is not None:
            n += 3      # reward, mppi partials, mppi combine
        return n

    def _enqueue(self):
        p = self.planner
        if self.general:
            from .planner import general_rollout
            with torch.no_grad():
                self.states.copy_(general_rollout(p, self.model_dy, self.s0, self.dens, self.attr, self.actions))
        else:
            ops.rollout_forward_raw(self._wpack, self.attr, self.dens, self.s0, self.actions, p.pusher,
                                    self.model_dy.adj_thresh, self.scratch, None, out=self.states)
        if self.goal_img is not None:
            N, T = self.N, self.T
            last = self.states[:, T - 1]
            lib = _lib.load()
            off = p.reward_offset()
            _lib.check(lib.pile_reward(_lib.ptr(last), self.rows, T * N * 3, N, _lib.ptr(self.goal_img),
                                       self.goal_img.shape[0], self.goal_img.shape[1], _lib.ptr(self.goal_coor),
                                       self.goal_coor.shape[0], _lib.host_floats(p.cam_params), off[0], off[1], 1,
                                       _lib.ptr(self.reward), None, ops._stream()), "pile_reward")
            _lib.check(lib.pile_mppi_partials(_lib.ptr(self.reward), _lib.ptr(self.actions), self.rows, T,
                                              float(self.reward_weight), _lib.ptr(self._parts), ops._stream()),
                       "pile_mppi_partials")
            _lib.check(lib.pile_mppi_combine(_lib.ptr(self._parts), self._parts.shape[0], T, _lib.ptr(self.record),
                                             ops._stream()), "pile_mppi_combine")

    def evaluate(self):
        """Roll `self.actions` out from `self.s0`; fills states, reward (last step) and the MPPI record."""
        wpack = self.model_dy.model.packed_weights(self.device)
        mode = _lib.load().pile_get_tensor_cores()
        # a captured graph keeps the weights buffer and the GEMM engine it was captured with
        if self._wpack is None or self._wpack.data_ptr() != wpack.data_ptr() or mode != self._mode:
            self._wpack, self._graph, self._mode = wpack, None, mode
        if not self.use_graph:
            self._enqueue()
            return
        if self._graph is None:
            self._enqueue()                     # warm-up outside capture (one-time kernel attribute setup)
            torch.cuda.synchronize(self.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._enqueue()
            self._graph = g
        self._graph.replay()

    def evaluate_host(self, actions_host, out_reward_host, out_record_host):
        """Reference-facing call with HOST buffers (pinned): H2D actions -> evaluate -> D2H reward + record."""
        self.actions.copy_(actions_host, non_blocking=True)
        self.evaluate()
        out_reward_host.copy_(self.reward, non_blocking=True)
        out_record_host.copy_(self.record, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()

    def mppi_mean(self):
        return (self.record[2:] / self.record[1]).view(self.T, 4)
