"""Generate tests/golden/golden_rgr_v1.npz by RUNNING THE REAL REFERENCE (build container only):
MPCResRgrNoPool (model/res_regressor.py:106-177) with its default initialisation under torch.manual_seed(5), on a
synthetic foreground mask and goal mask: the 6-plane input `infer_param` builds, the raw network output and the
returned particle count.

    python tests/golden/make_golden_rgr.py
"""
import importlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
import ref_harness  # noqa: E402
from dyn_res_pile_manip_b200 import synthetic  # noqa: E402

SEED = 5


def main():
    ref_harness.load_reference()
    rr = importlib.import_module("model.res_regressor")
    cfg = {"train_res_cls": {"state_h": 224, "state_w": 224, "res_dim": 6}}
    torch.manual_seed(SEED)
    net = rr.MPCResRgrNoPool(cfg)
    env = synthetic.FakeEnv()
    st, _ = synthetic.make_pile_batch(1, 200, seed=4)
    obs = synthetic.render_observation(st[0], env)
    fg = (obs[..., -1] / env.global_scale < 0.599 / 0.8).astype(np.float32)
    goal = (synthetic.make_goal("tee") < 0.5).astype(np.float32)
    grabbed = {}
    fwd = net.forward

    def spy(x):
        grabbed["x"] = x.detach().cpu().numpy().copy()
        out = fwd(x.cpu())
        grabbed["y"] = out.detach().numpy().copy()
        return out
    net.forward = spy
    cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self          # infer_param calls .cuda() unconditionally (:175)
    try:
        n = net.infer_param(fg, goal)
    finally:
        torch.Tensor.cuda = cuda
    path = os.path.join(HERE, "golden_rgr_v1.npz")
    np.savez_compressed(path, fg=fg.astype(np.uint8), goal=goal.astype(np.uint8), x=grabbed["x"][0], y=grabbed["y"],
                        particle_num=np.asarray(n), seed=np.asarray(SEED))
    print("wrote", path, "output", grabbed["y"], "->", n)


if __name__ == "__main__":
    main()
