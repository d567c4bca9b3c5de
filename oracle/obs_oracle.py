"""CPU ORACLE for the observation -> particles step.  TEST INFRASTRUCTURE ONLY (same rules as pile_oracle.py: only
tests/, __graft_entry__.smoke() and bench.py's CPU baseline legs may import it).

numpy restatement of the reference functions, each in the reference's own dtypes:
  depth2fgpcd      utils.py:491-506      (pure numpy in the reference: PINNED by tests/golden/golden_obs_v1.npz)
  recenter         utils.py:468-477      (pure numpy in the reference: PINNED by the same fixture)
  fps              utils.py:423-437      sampler = dgl.geometry.farthest_point_sampler (third party, env.yaml lists
                                         `dgl` without a version; absent from /root/reference and from this image):
                                         restated from its published algorithm - start at the given index, keep the
                                         minimum SQUARED distance to the picked set in float32, take the first argmax.
                                         PARITY UNPINNED for the sampler itself; utils.fps's own numpy part
                                         (particle_r) is covered by the fixture through fps_np's identical code.
  downsample_pcd   utils.py:533-544      open3d PointCloud::VoxelDownSample (third party, env.yaml lists `open3d`
                                         without a version; absent here): restated - voxel_min_bound = min_bound -
                                         voxel/2, index = floor((p - voxel_min_bound) / voxel), output = mean of the
                                         points of each voxel accumulated in input order.  PARITY UNPINNED; open3d
                                         emits the voxels in hash-map order, this restatement in ascending
                                         (ix, iy, iz), and the tests compare as sets where order is unspecified.
  obs2ptcl_fixed_num_batch  env/flex_env.py:933-951
"""
import numpy as np


def depth2fgpcd(depth, mask, cam_params):
    mask = np.logical_and(mask, depth > 0)
    fgpcd = np.zeros((mask.sum(), 3))
    fx, fy, cx, cy = cam_params
    pos_x, pos_y = np.meshgrid(np.arange(depth.shape[1]), np.arange(depth.shape[0]))
    pos_x, pos_y = pos_x[mask], pos_y[mask]
    fgpcd[:, 0] = (pos_x - cx) * depth[mask] / fx
    fgpcd[:, 1] = (pos_y - cy) * depth[mask] / fy
    fgpcd[:, 2] = depth[mask]
    return fgpcd


def voxel_down_sample(pcd, voxel_size):
    pcd = np.asarray(pcd, dtype=np.float64)
    lo = pcd.min(axis=0) - voxel_size * 0.5
    idx = np.floor((pcd - lo) / voxel_size).astype(np.int64)
    key = (idx[:, 0] << 42) | (idx[:, 1] << 21) | idx[:, 2]
    order = np.argsort(key, kind="stable")
    ks = key[order]
    heads = np.flatnonzero(np.r_[True, ks[1:] != ks[:-1]])
    out = np.empty((len(heads), 3))
    ends = np.r_[heads[1:], len(ks)]
    for s, (a, b) in enumerate(zip(heads, ends)):
        acc = np.zeros(3)
        for j in order[a:b]:              # input order inside the voxel, left-to-right sum like open3d's accumulator
            acc = acc + pcd[j]
        out[s] = acc / float(b - a)
    return out


def farthest_point_sampler(pcd32, count, init_idx):
    """float32 cloud [m,3], squared distances, first index on ties -> indices [count]."""
    p = np.asarray(pcd32, dtype=np.float32)
    idx = np.empty(count, dtype=np.int64)
    idx[0] = init_idx
    d = p - p[init_idx]
    gap = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
    for n in range(1, count):
        idx[n] = int(gap.argmax())
        d = p - p[idx[n]]
        gap = np.minimum(gap, (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2])
    return idx


def fps(pcd, particle_num, init_idx):
    pcd_fps = np.asarray(pcd, dtype=np.float32)[farthest_point_sampler(pcd, particle_num, init_idx)]
    dist = np.linalg.norm(pcd[:, None] - pcd_fps[None, :], axis=-1).min(axis=1)
    return pcd_fps, dist.max()


def recenter(pcd, sampled_pcd, r=0.02):
    dist = np.linalg.norm(pcd[:, None, :] - sampled_pcd[None, :, :], axis=2)
    out = np.zeros_like(sampled_pcd)
    for i in range(sampled_pcd.shape[0]):
        out[i] = pcd[dist[:, i] < r].mean(axis=0)
    return out


def obs2ptcl_fixed_num_batch(obs, particle_num, batch_size, cam_params, global_scale, init_idx):
    depth = obs[..., -1] / global_scale
    out = np.zeros((batch_size, particle_num, 3))
    rad = np.zeros((batch_size,))
    fgpcd = voxel_down_sample(depth2fgpcd(depth, depth < 0.599 / 0.8, cam_params), 0.01)   # same for every i
    for i in range(batch_size):
        sampled, particle_r = fps(fgpcd, particle_num, int(init_idx[i]))
        out[i] = recenter(fgpcd, sampled, r=min(0.02, 0.5 * particle_r))
        rad[i] = particle_r
    return out, rad, fgpcd
