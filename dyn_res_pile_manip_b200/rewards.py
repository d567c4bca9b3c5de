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
    """shaped goal image on the device, keyed by the goal tensor's storage + version."""

    def __init__(self):
        self._key = None
        self._img = None

    def shaped(self, goal):
        key = (goal.data_ptr(), goal._version, tuple(goal.shape), str(goal.device))
        if key != self._key:
            self._img = shape_goal_image(goal)
            self._key = key
        return self._img


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
