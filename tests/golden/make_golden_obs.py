"""Generate tests/golden/golden_obs_v1.npz by RUNNING THE REAL REFERENCE helpers (build container only).

Covers the pure-numpy parts of the observation -> particles step: utils.depth2fgpcd (utils.py:491-506),
utils.recenter (:468-477) and utils.fps_np (:451-466, whose distance/covering-radius code is the numpy part of
utils.fps).  utils.downsample_pcd (open3d) and dgl's farthest_point_sampler are third-party and absent here; see the
header of oracle/obs_oracle.py.

    python tests/golden/make_golden_obs.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
import ref_harness  # noqa: E402
from dyn_res_pile_manip_b200 import synthetic  # noqa: E402


def main():
    ref = ref_harness.load_reference()
    U = ref["utils"]
    env = synthetic.FakeEnv()
    out = {}
    # a 96x96 crop-sized scene keeps the fixture small: 12 particles, same camera model
    rng = np.random.RandomState(7)
    size = 96
    cam = [110.0, 110.0, size / 2.0, size / 2.0]

    class Cam(synthetic.FakeEnv):
        def get_cam_params(self):
            return cam

    state = np.stack([rng.uniform(-0.25, 0.25, 12), rng.uniform(-0.25, 0.25, 12), np.full(12, 0.74)], 1)
    obs = synthetic.render_observation(state, Cam(), particle_radius=0.03, size=size)
    depth = obs[..., -1] / env.global_scale
    mask = depth < 0.599 / 0.8
    fg = U.depth2fgpcd(depth, mask, cam)
    out["obs"] = obs
    out["cam"] = np.asarray(cam)
    out["global_scale"] = np.asarray(env.global_scale)
    out["fgpcd"] = fg
    # recenter + fps_np on a voxel-sized cloud
    cloud = np.stack([rng.uniform(-0.1, 0.1, 400), rng.uniform(-0.1, 0.1, 400), 0.74 - rng.uniform(0, 0.01, 400)], 1)
    picks32 = cloud[rng.choice(400, 20, replace=False)].astype(np.float32)
    out["cloud"] = cloud
    out["picks32"] = picks32
    out["recenter_r02"] = U.recenter(cloud, picks32, r=0.02)
    out["recenter_r007"] = U.recenter(cloud, picks32, r=0.007)
    chosen, radius = U.fps_np(cloud, 25, 3)
    out["fps_np_chosen"] = chosen
    out["fps_np_radius"] = np.asarray(radius)
    path = os.path.join(HERE, "golden_obs_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: np.asarray(v).shape for k, v in out.items()})


if __name__ == "__main__":
    main()
