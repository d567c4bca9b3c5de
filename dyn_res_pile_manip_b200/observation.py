"""Observation -> particles on the GPU: the planner-side work of one MPC step before the rollout.

Mirrors `FlexEnv.obs2ptcl_fixed_num_batch` / `obs2ptcl_fixed_num` (reference env/flex_env.py:910-951) and the
helpers they call (utils.depth2fgpcd :491-506, downsample_pcd :533-544, fps :423-437, recenter :468-477) with the
same names, argument meaning and return types (numpy in, numpy out).  Everything between the depth image and the
particle set stays on the device (kernels in csrc/obs.cu, farthest-point sampling in csrc/reward.cu); the
reference runs it on the host with open3d + dgl, 30 times per MPC step.

Two things the reference leaves unspecified are made explicit here: the order of the voxel-downsampled cloud
(open3d iterates a hash map; here ascending voxel index) and the sampler's start points (dgl draws them at random;
here `init_idx` or a seeded numpy generator).  There is no CPU fallback.
"""
import numpy as np
import torch

from . import _lib, ops

VOXEL_SIZE = 0.01              # flex_env.py:947
FG_MAX_DEPTH = 0.599 / 0.8     # flex_env.py:946
RECENTER_CAP = 0.02            # flex_env.py:949


def _dev(device):
    return torch.device(device if device is not None else "cuda")


def depth2fgpcd(depth, cam_params, max_depth=FG_MAX_DEPTH, device=None):
    """depth [H,W] float32 (numpy or CUDA tensor) -> foreground cloud [n,3] float64 CUDA tensor in pixel order
    (utils.depth2fgpcd with mask = depth < max_depth, and depth > 0 as the function itself adds)."""
    dev = _dev(device)
    d = torch.as_tensor(depth, dtype=torch.float32).to(dev).contiguous()
    H, W = d.shape
    lib = _lib.load()
    counts = torch.empty(lib.pile_depth_counts_len(H, W), dtype=torch.int32, device=dev)
    pts = torch.empty(H * W, 3, dtype=torch.float64, device=dev)
    n = torch.zeros(1, dtype=torch.int32, device=dev)
    _lib.check(lib.pile_depth_to_points(_lib.ptr(d), H, W, _lib.host_doubles(cam_params), float(np.float32(max_depth)),
                                        _lib.ptr(pts), H * W, _lib.ptr(n), _lib.ptr(counts), ops._stream()),
               "pile_depth_to_points")
    return pts[: int(n.item())]


def downsample_pcd(pcd, voxel_size=VOXEL_SIZE):
    """pcd [n,3] float64 CUDA -> [m,3] float64 CUDA, one point per occupied voxel (ascending voxel index)."""
    pcd = pcd.contiguous()
    n = pcd.shape[0]
    if n == 0:
        return pcd
    lib = _lib.load()
    ws = torch.empty(lib.pile_voxel_downsample_bytes(n), dtype=torch.uint8, device=pcd.device)
    out = torch.empty(n, 3, dtype=torch.float64, device=pcd.device)
    m = torch.zeros(1, dtype=torch.int32, device=pcd.device)
    _lib.check(lib.pile_voxel_downsample(_lib.ptr(pcd), n, float(voxel_size), _lib.ptr(out), _lib.ptr(m), _lib.ptr(ws),
                                         ops._stream()), "pile_voxel_downsample")
    return out[: int(m.item())]


def fps_batch(pcd, particle_num, init_idx):
    """utils.fps for a batch of start indices: pcd [m,3] float64 CUDA, init_idx [S] ints ->
    (picks [S,N,3] float32 CUDA, particle_r [S] float64 CUDA).  The sampler runs on the float32 copy of the cloud and
    compares squared distances like dgl's; particle_r is measured on the float64 cloud like utils.fps."""
    dev = pcd.device
    m = pcd.shape[0]
    init = torch.as_tensor(np.asarray(init_idx), dtype=torch.int32).to(dev).contiguous()
    S = init.numel()
    p32 = pcd.to(torch.float32).contiguous()
    lib = _lib.load()
    gap = torch.empty(S, m, dtype=torch.float32, device=dev)
    idx = torch.empty(S, particle_num, dtype=torch.int32, device=dev)
    picks = torch.empty(S, particle_num, 3, dtype=torch.float32, device=dev)
    _lib.check(lib.pile_fps_sets(_lib.ptr(p32), 1, S, m, 3, int(particle_num), _lib.ptr(init), 1, _lib.ptr(gap),
                                 _lib.ptr(idx), _lib.ptr(picks), None, ops._stream()), "pile_fps_sets")
    rad = torch.empty(S, dtype=torch.float64, device=dev)
    _lib.check(lib.pile_cover_radius(_lib.ptr(pcd), m, _lib.ptr(picks), S, int(particle_num), _lib.ptr(rad),
                                     ops._stream()), "pile_cover_radius")
    return picks, rad, idx


def recenter_batch(pcd, picks, particle_r, r_cap=RECENTER_CAP, r_scale=0.5):
    """utils.recenter with r = min(r_cap, r_scale * particle_r) per set -> [S,N,3] float32 CUDA."""
    S, N, _ = picks.shape
    out = torch.empty_like(picks)
    _lib.check(_lib.load().pile_recenter(_lib.ptr(pcd), pcd.shape[0], _lib.ptr(picks), S, N, _lib.ptr(particle_r),
                                         float(r_cap), float(r_scale), _lib.ptr(out), ops._stream()), "pile_recenter")
    return out


def obs2ptcl_fixed_num_batch(obs, particle_num, batch_size, cam_params, global_scale, init_idx=None, seed=None,
                             device=None, return_device=False):
    """FlexEnv.obs2ptcl_fixed_num_batch (flex_env.py:933-951): obs [H,W,5] (RGB, -, depth*global_scale) ->
    (batch_sampled_ptcl [batch_size, particle_num, 3], batch_particle_r [batch_size]) as float64 numpy arrays
    (CUDA tensors with return_device=True).  `init_idx` [batch_size]: start index of each sampler run in the
    downsampled cloud (default: drawn from numpy's RandomState(seed), where the reference lets dgl draw them)."""
    dev = _dev(device)
    if isinstance(obs, np.ndarray):
        assert obs.shape[-1] == 5
        depth = torch.from_numpy(np.ascontiguousarray(obs[..., -1])).to(dev, torch.float32)
    else:
        depth = obs[..., -1].to(dev, torch.float32)
    # tensor / tensor: torch divides by a python scalar as a multiplication with its reciprocal, numpy does not
    depth = depth / torch.full((), float(global_scale), dtype=torch.float32, device=dev)
    fgpcd = downsample_pcd(depth2fgpcd(depth, cam_params, device=dev), VOXEL_SIZE)
    m = fgpcd.shape[0]
    if m < particle_num:
        raise ValueError("foreground cloud has %d points after downsampling, fewer than particle_num=%d" % (m, particle_num))
    if init_idx is None:
        init_idx = np.random.RandomState(seed).randint(m, size=batch_size)
    picks, rad, _ = fps_batch(fgpcd, particle_num, init_idx)
    out = recenter_batch(fgpcd, picks, rad)
    if return_device:
        return out, rad
    return out.to(torch.float64).cpu().numpy(), rad.cpu().numpy()


def obs2ptcl_fixed_num(obs, particle_num, cam_params, global_scale, init_idx=None, seed=None, device=None):
    """FlexEnv.obs2ptcl_fixed_num (flex_env.py:910-931): one sampler run -> (sampled_ptcl [N,3], particle_r)."""
    pts, r = obs2ptcl_fixed_num_batch(obs, particle_num, 1, cam_params, global_scale,
                                      None if init_idx is None else [init_idx], seed, device)
    return pts[0], float(r[0])


def shift_warm_start(action_seq_mpc_init, action_full, n_look_ahead):
    """Warm start of the next MPC step, the reference's expression verbatim (flex_env.py:1112-1113).  With the
    shipped settings the initial sequence has length 1 and it is returned unchanged; for longer sequences the
    reference concatenates `action_full[1:]` ([traj-1, 4]) with a [steps, traj, 4] array, which numpy rejects -
    that behaviour (a ValueError) is kept rather than guessed around."""
    if action_seq_mpc_init.shape[0] > 1:
        return np.concatenate((action_full[1:], action_seq_mpc_init[n_look_ahead:]), axis=0)
    return action_seq_mpc_init
