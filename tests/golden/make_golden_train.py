"""Generate tests/golden/golden_train_v1.npz by RUNNING THE REAL REFERENCE (build container only): the training
objective of train/train_gnn_dyn.py:150-192 (multi-step rollout, per-sample MSE over the valid particles, mean over
n_rollout * B) and its autograd gradients w.r.t. all 18 weight tensors - the pinned target for the training path
(wgrad + particle_nums masking).  Weights: the seed-0 model of make_golden.py
(golden_v1.npz, keys w/...).

    python tests/golden/make_golden_train.py
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
import ref_harness  # noqa: E402
from dyn_res_pile_manip_b200 import synthetic  # noqa: E402


def main():
    ref = ref_harness.load_reference()
    torch.set_num_threads(4)
    cfg = synthetic.default_config()
    torch.manual_seed(0)
    model = ref["gnn_dyn"].PropNetDiffDenModel(cfg, False)          # same weights as golden_v1.npz
    B, N, n_roll = 3, 40, 2
    rng = np.random.RandomState(21)
    states0, dens = synthetic.make_pile_batch(B, N, seed=21)
    particle_nums = np.array([40, 33, 25], dtype=np.int64)          # padded variable-N batch (collate :20-43)
    states = np.stack([states0 + rng.normal(0, 0.004, states0.shape).astype(np.float32) * k for k in range(n_roll + 1)], 1)
    states_delta = (rng.normal(0, 0.01, (B, n_roll, N, 3)) * (rng.uniform(size=(B, n_roll, N, 1)) < 0.3)).astype(np.float32)
    attrs = np.zeros((B, n_roll + 1, N), dtype=np.float32)
    st, sd, at = torch.tensor(states), torch.tensor(states_delta), torch.tensor(attrs)
    pd, pn = torch.tensor(dens), torch.tensor(particle_nums)

    loss = 0.
    s_cur, a_cur = st[:, 0], at[:, 0]
    for step in range(n_roll):
        s_nxt = st[:, step + 1]
        s_pred = model.predict_one_step(a_cur, s_cur, sd[:, step], pd, pn)
        for j in range(B):
            loss = loss + F.mse_loss(s_pred[j, :particle_nums[j]], s_nxt[j, :particle_nums[j]])
        s_cur = s_pred
    loss = loss / (n_roll * B)
    loss.backward()
    out = {"states": states, "states_delta": states_delta, "attrs": attrs, "dens": dens, "particle_nums": particle_nums,
           "loss": np.asarray(loss.item())}
    for k, p in model.named_parameters():
        out["g/" + k] = p.grad.numpy().copy()
    path = os.path.join(HERE, "golden_train_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "loss", loss.item(), "params", sum(1 for _ in model.parameters()))


if __name__ == "__main__":
    main()
