"""Synthetic piles, goals and a simulator-free stand-in for FlexEnv.

Nothing here touches the GPU.  `FakeEnv` carries exactly the attributes `PlannerGD`
reads from the real environment (reference planners.py:32-45, 152-155): camera
intrinsics of the FleX 45-degree camera (PyFleX/bindings/pyflex.cpp:3501-3518,
env/flex_env.py:1135-1142), the view matrix of the top-down camera 18 units above the
table (env/flex_env.py:194-200, pyflex.cpp:3484-3499), the 720x720 screen and the
workspace box (env/flex_env.py:454-458).
"""
import math

import numpy as np

SCREEN = 720
GLOBAL_SCALE = 24.0
WKSPC_W = 5.0
PILE_HALF_EXTENT = 0.15
PILE_Z = 0.74


def default_config(nf_effect=64, adj_thresh=0.08):
    """The keys of config/mpc/config.yaml that the hot path reads (SURVEY.md §5 'Config')."""
    return {
        "dataset": {"global_scale": GLOBAL_SCALE, "wkspc_w": WKSPC_W},
        "mpc": {"sigma": 0.3, "mppi": {"beta_filter": 0.7, "reward_weight": 0.1},
                "gd": {"beta_filter": 0.7, "lr": 0.05},
                "n_look_ahead": 1, "n_sample": 50, "n_update_iter": 200, "time_lim": 2000},
        "train": {"n_history": 1,
                  "particle": {"nf_effect": nf_effect, "adj_thresh": adj_thresh, "add_delta": False}},
    }


class FakeEnv:
    """Camera + workspace constants of the FleX scene, no simulator behind it."""
    is_real = False
    screenHeight = SCREEN
    screenWidth = SCREEN

    def __init__(self, global_scale=GLOBAL_SCALE, wkspc_w=WKSPC_W):
        self.global_scale = global_scale
        self.wkspc_w = wkspc_w
        self.cvx_region = np.array([[-wkspc_w, wkspc_w, -wkspc_w, wkspc_w]], dtype=np.float64)
        f = (SCREEN / 2.0) / math.tan(math.radians(45.0) / 2.0)
        self._cam = [f, f, SCREEN / 2.0, SCREEN / 2.0]
        cam_height = 6.0 * global_scale / 8.0
        rot = np.array([[1, 0, 0, 0], [0, 0, -1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=np.float64)
        shift = np.eye(4)
        shift[1, 3] = -cam_height
        self._view = (rot @ shift).astype(np.float32)

    def get_cam_params(self):
        return list(self._cam)

    def get_cam_extrinsics(self):
        return self._view.copy()


def fps_np(pts, count, init_idx=0):
    """Farthest-point sampling, same contract as reference utils.fps_np (utils.py:451-466):
    start from `init_idx`, repeatedly take the point farthest from the chosen set.
    -> (chosen [count,c], covering radius)."""
    pts = np.asarray(pts)
    if init_idx < 0:
        init_idx = np.random.randint(pts.shape[0])
    chosen = np.empty((count,) + pts.shape[1:], dtype=pts.dtype)
    chosen[0] = pts[init_idx]
    gap = np.linalg.norm(pts - chosen[0], axis=1)
    for n in range(1, count):
        chosen[n] = pts[gap.argmax()]
        gap = np.minimum(gap, np.linalg.norm(pts - chosen[n], axis=1))
    return chosen, gap.max()


def make_pile(n_particles, seed=0, candidates=20000):
    """Seed-fixed FPS-sampled pile in the camera frame -> (state [N,3] f32, density float)."""
    rng = np.random.RandomState(seed)
    cand = rng.uniform(-PILE_HALF_EXTENT, PILE_HALF_EXTENT, size=(candidates, 2))
    xy, radius = fps_np(cand, n_particles, 0)
    state = np.concatenate([xy, np.full((n_particles, 1), PILE_Z)], axis=1).astype(np.float32)
    return state, float(1.0 / (radius * radius))


def make_pile_batch(n_batch, n_particles, seed=0, jitter=1e-3):
    """[n_batch,N,3] state variants of one pile (stand-in for the 30 FPS resamplings of
    env/flex_env.py:1020) + per-variant densities [n_batch]."""
    base, dens = make_pile(n_particles, seed)
    rng = np.random.RandomState(seed + 1000)
    states = np.repeat(base[None], n_batch, axis=0)
    if n_batch > 1 and jitter > 0:
        noise = rng.normal(0, jitter, size=states.shape).astype(np.float32)
        noise[..., 2] = 0
        noise[0] = 0
        states = states + noise
    d = dens * (1.0 + 0.05 * rng.uniform(-1, 1, size=n_batch))
    d[0] = dens
    return states.astype(np.float32), d.astype(np.float32)


def make_goal(kind="bar", size=SCREEN):
    """Synthetic goal distance image [size,size] f32 (0 inside the target shape, L2 pixel
    distance outside, capped at 1e4 like utils.gen_goal_shape, utils.py:566-579)."""
    import cv2
    mask = np.zeros((size, size), dtype=np.uint8)
    c = size // 2
    if kind == "bar":          # an 'I'-like vertical bar
        mask[c - size // 4:c + size // 4, c - size // 18:c + size // 18] = 1
    elif kind == "disc":
        yy, xx = np.mgrid[0:size, 0:size]
        mask[(yy - c) ** 2 + (xx - c) ** 2 < (size // 8) ** 2] = 1
    elif kind == "tee":
        mask[c - size // 5:c - size // 8, c - size // 5:c + size // 5] = 1
        mask[c - size // 8:c + size // 4, c - size // 24:c + size // 24] = 1
    else:
        raise ValueError(kind)
    return np.minimum(cv2.distanceTransform(1 - mask, cv2.DIST_L2, 5), 1e4).astype(np.float32)


def random_actions(n_rows, horizon, seed=0, lim=4.0):
    rng = np.random.RandomState(seed + 7)
    return rng.uniform(-lim, lim, size=(n_rows, horizon, 4)).astype(np.float32)


def render_observation(state, env=None, particle_radius=0.008, table_depth=0.75, size=SCREEN):
    """Synthetic top-down RGB-D observation of a pile in the layout FlexEnv.render returns ([H,W,5] float32: RGB,
    unused, depth * global_scale): every particle (camera-frame x, y, z) is drawn as a sphere above a flat table at
    `table_depth` (reference scenes: table at 0.75, foreground test depth < 0.599/0.8, flex_env.py:927)."""
    env = env if env is not None else FakeEnv()
    fx, fy, cx, cy = env.get_cam_params()
    depth = np.full((size, size), table_depth, dtype=np.float64)
    for x, y, z in np.asarray(state, dtype=np.float64):
        u0, v0 = x * fx / z + cx, y * fy / z + cy
        rp = particle_radius * fx / z
        ua, ub = max(int(u0 - rp) - 1, 0), min(int(u0 + rp) + 2, size)
        va, vb = max(int(v0 - rp) - 1, 0), min(int(v0 + rp) + 2, size)
        if ua >= ub or va >= vb:
            continue
        uu, vv = np.meshgrid(np.arange(ua, ub), np.arange(va, vb))
        rho2 = ((uu - u0) / fx * z) ** 2 + ((vv - v0) / fy * z) ** 2
        inside = rho2 < particle_radius ** 2
        surf = z - np.sqrt(np.maximum(particle_radius ** 2 - rho2, 0.0))
        patch = depth[va:vb, ua:ub]
        patch[inside] = np.minimum(patch[inside], surf[inside])
    obs = np.zeros((size, size, 5), dtype=np.float32)
    obs[..., :3] = 255.0
    obs[depth < table_depth, :3] = 128.0
    obs[..., 4] = (depth * env.global_scale).astype(np.float32)
    return obs
