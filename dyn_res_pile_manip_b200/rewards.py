"""Particle-space target-shape reward (reference env/flex_rewards.py:156-214) on libpilegnn.

`config_reward_ptcl(state, goal, cam_params, goal_coor, normalize=True, offset=(0., 0.))`
keeps the reference signature.  The only host work is the one-off shaping of the goal
image with OpenCV's distance transform (flex_rewards.py:171-176), cached per goal instead
of being recomputed (and tiled B times) on every call.
"""
import numpy as np
import torch

from . import ops


class GoalCache:
    """Shaped goal image on the device, keyed by the goal's CONTENT.

    A key made of (data_ptr, _version) is not enough: the planner builds a fresh goal tensor per call and the
    caching allocator hands the next same-size goal the same address, so a different goal would hit the stale
    entry.  The cache therefore keeps its own copy of the keyed goal; a lookup is a hit when it is handed the very
    same tensor object at the same version (no sync), or a tensor that compares equal to the kept copy."""

    def __init__(self):
        self._ref = None        # the tensor object the entry was built from (kept alive: its address cannot be reused)
        self._version = None
        self._copy = None       # private copy of its contents at that time
        self._img = None
        self._np_key = None

    def shaped(self, goal):
        if self._img is not None and self._copy is not None:
            if goal is self._ref and goal._version == self._version:
                return self._img
            if goal.shape == self._copy.shape and goal.device == self._copy.device and goal.dtype == self._copy.dtype \
                    and torch.equal(goal, self._copy):
                self._ref, self._version = goal, goal._version
                return self._img
        self._img = shape_goal_image(goal)
        self._ref, self._version, self._copy = goal, goal._version, goal.detach().clone()
        self._np_key = None
        return self._img

    def shaped_np(self, goal_np, goal_tensor, content_hash=None, device=None):
        """Planner entry: the goal arrives as a numpy array, so the key is a hash of its bytes (as
        PlannerGD.goal_coordinates does; `content_hash` = goal_content_hash(goal_np) when the caller already has it);
        `goal_tensor` is the same image already on the device, or None with `device` given: the image is then copied
        to the device only on a miss (2 MB for a 720 x 720 goal, every call otherwise)."""
        g = np.ascontiguousarray(goal_np)
        dev = goal_tensor.device if goal_tensor is not None else torch.device(device)
        key = (g.shape, str(g.dtype), str(dev), goal_content_hash(g) if content_hash is None else content_hash)
        if self._img is None or key != self._np_key:
            if goal_tensor is None:
                goal_tensor = torch.as_tensor(g, dtype=torch.float32).to(dev)
            self._img = shape_goal_image(goal_tensor)
            self._np_key = key
            self._ref, self._version, self._copy = goal_tensor, goal_tensor._version, goal_tensor.detach().clone()
        return self._img


def goal_content_hash(goal_np):
    """Hash of the goal image's bytes: the cache key of everything derived from a goal (0.45 ms for 720 x 720 floats, so a
    planner call computes it once and hands it to both users)."""
    return hash(np.ascontiguousarray(goal_np).tobytes())


def shape_goal_image(goal):
    """goal' = goal - DT(goal < 0.5) - min(...)   (flex_rewards.py:171-176)."""
    import cv2
    g = goal.detach().cpu().numpy().astype(np.float32)
    inside = cv2.distanceTransform((g < 0.5).astype(np.uint8), cv2.DIST_L2, 5)
    g = g - inside
    g = g - g.min()
    return torch.from_numpy(np.ascontiguousarray(g)).to(device=goal.device, dtype=torch.float32)


class _RewardFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, state, img, goal_coor, cam_params, offset, normalize):
        st = ops._f32(state.detach())
        ops._require_cuda(st, "state")
        B, N, _ = st.shape
        need = state.requires_grad
        r, arg = ops.reward_raw(st, B, N * 3, N, img, goal_coor, cam_params, offset, normalize, want_argmin=need)
        ctx.saved = (st, img, goal_coor, cam_params, offset, normalize, arg)
        return r

    @staticmethod
    def backward(ctx, g):
        st, img, goal_coor, cam_params, offset, normalize, arg = ctx.saved
        B, N, _ = st.shape
        g_state = torch.empty_like(st)
        ops.reward_backward_raw(st, B, N * 3, N, img, goal_coor, cam_params, offset, normalize, ops._f32(g), arg,
                                g_state, N * 3, False)
        return g_state, None, None, None, None, None


def config_reward_ptcl(state, goal, cam_params, goal_coor, normalize=True, offset=(0., 0.), cache=None):
    """state [B,N,3], goal [H,W], cam_params (fx,fy,cx,cy), goal_coor [M,2]=(col,row) -> [B]."""
    B, N, _ = state.shape
    assert state.shape[2] == 3
    assert type(state) == torch.Tensor
    assert type(goal) == torch.Tensor
    img = cache.shaped(goal) if cache is not None else shape_goal_image(goal)
    coor = ops._f32(goal_coor, state.device)
    cam = [float(v) for v in cam_params]
    return _RewardFn.apply(state, img, coor, cam, (float(offset[0]), float(offset[1])), bool(normalize))
