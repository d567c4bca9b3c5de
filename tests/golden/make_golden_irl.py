"""Generate tests/golden/golden_irl_v1.npz by RUNNING THE REAL REFERENCE (build container only):
PlannerGD.gen_s_delta_irl (planners.py:259-300), the real-robot pusher parametrisation behind env.is_real
(SURVEY section 8f rank 4).  The real environment class is not part of the reference tree; the three attributes the
function reads (wkspc_center_x, wkspc_center_y, s2r_scale) are set on the synthetic environment.

    python tests/golden/make_golden_irl.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
import ref_harness  # noqa: E402
from dyn_res_pile_manip_b200 import synthetic  # noqa: E402

IRL_ENV = {"wkspc_center_x": 0.03, "wkspc_center_y": -0.02, "s2r_scale": 10.0}


def main():
    ref = ref_harness.load_reference()
    cfg, env = synthetic.default_config(), synthetic.FakeEnv()
    for k, v in IRL_ENV.items():
        setattr(env, k, v)
    planner = ref["planners"].PlannerGD(cfg, env)
    N = 80
    planner.particle_num = N
    states, _ = synthetic.make_pile_batch(5, N, seed=11)
    states = states.copy()
    states[..., 2] = 0.88                      # the real camera frame keeps particles at the pusher height
    states[..., 0] += IRL_ENV["wkspc_center_x"]
    states[..., 1] += IRL_ENV["wkspc_center_y"]
    act = np.array([[-1.2, 0.3, 1.4, -0.2], [0.2, -1.5, -0.1, 1.3], [-1.0, -1.0, 1.0, 1.0], [1.3, 0.1, -1.2, 0.0],
                    [0.0, 0.0, 0.5, 0.5]], dtype=np.float32)
    s = torch.tensor(states, requires_grad=True)
    a = torch.tensor(act, requires_grad=True)
    out = planner.gen_s_delta_irl(s, a)
    w = torch.tensor(np.random.RandomState(0).normal(size=out.shape).astype(np.float32))
    (out * w).sum().backward()
    path = os.path.join(HERE, "golden_irl_v1.npz")
    np.savez_compressed(path, s_cur=states, action=act, s_delta=out.detach().numpy(), weight=w.numpy(),
                        g_s_cur=s.grad.numpy(), g_action=a.grad.numpy(),
                        **{k: np.asarray(v) for k, v in IRL_ENV.items()})
    print("wrote", path, "moved particles:", int((out.detach().abs().sum(-1) > 0).sum()))


if __name__ == "__main__":
    main()
