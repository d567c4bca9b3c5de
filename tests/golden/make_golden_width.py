"""Generate tests/golden/golden_width_v1.npz by RUNNING THE REAL REFERENCE (build container only) at a hidden width
other than the shipped 64 (model/gnn_dyn.py:119 reads nf_effect from the config): nf_effect = 96 -- not a multiple of
64, so the general-width engine (csrc/general.cu) pads it to 128.  Contents: the seed-0 weights of that model, the
training objective of train/train_gnn_dyn.py:150-192 on a padded variable-N batch with its autograd gradients of all 18
tensors, and d(loss)/d(states[:, 0]), d(loss)/d(states_delta) -- the pinned target of tests/test_gpu_general_width.py.

    python tests/golden/make_golden_width.py
"""
import copy
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

NF = 96


def main():
    ref = ref_harness.load_reference()
    torch.set_num_threads(4)
    cfg = copy.deepcopy(synthetic.default_config())
    cfg['train']['particle']['nf_effect'] = NF
    torch.manual_seed(0)
    model = ref["gnn_dyn"].PropNetDiffDenModel(cfg, False)
    # A ReLU whose pre-activation is within a few fp32 ulps of zero has an implementation-defined mask (summation
    # order), and in a batch this small one flipped mask moves the weight gradients by 1e-4 .. 1e-2: such inputs pin
    # nothing.  Record the smallest |pre-activation| of every ReLU and take the first data seed without a near-tie.
    tiny = []
    for mod in model.modules():
        if isinstance(mod, torch.nn.ReLU):
            mod.register_forward_hook(lambda m, inp, out: tiny.append(float(inp[0].detach().abs().min())))
    for seed in range(23, 60):
        tiny.clear()
        out = run(model, seed)
        if min(tiny) > 4e-7:
            break
        print("seed", seed, "has a near-tie ReLU (min |pre-activation| %.1e): skipped" % min(tiny))
    out["data_seed"] = np.asarray(seed)
    out["min_abs_preactivation"] = np.asarray(min(tiny))
    path = os.path.join(HERE, "golden_width_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "loss", float(out["loss"]), "seed", seed, "min |pre-act| %.1e" % min(tiny), "bytes", os.path.getsize(path))


def run(model, seed):
    model.zero_grad()
    B, N, n_roll = 3, 40, 2
    rng = np.random.RandomState(seed)
    states0, dens = synthetic.make_pile_batch(B, N, seed=seed)
    particle_nums = np.array([40, 31, 26], dtype=np.int64)
    states = np.stack([states0 + rng.normal(0, 0.004, states0.shape).astype(np.float32) * k for k in range(n_roll + 1)], 1)
    states_delta = (rng.normal(0, 0.01, (B, n_roll, N, 3)) * (rng.uniform(size=(B, n_roll, N, 1)) < 0.3)).astype(np.float32)
    attrs = np.zeros((B, n_roll + 1, N), dtype=np.float32)
    st, at = torch.tensor(states), torch.tensor(attrs)
    sd = torch.tensor(states_delta, requires_grad=True)
    s0 = st[:, 0].clone().requires_grad_(True)
    pd, pn = torch.tensor(dens), torch.tensor(particle_nums)

    loss = 0.
    s_cur, a_cur = s0, at[:, 0]
    preds = []
    for step in range(n_roll):
        s_nxt = st[:, step + 1]
        s_pred = model.predict_one_step(a_cur, s_cur, sd[:, step], pd, pn)
        preds.append(s_pred.detach().numpy().copy())
        for j in range(B):
            loss = loss + F.mse_loss(s_pred[j, :particle_nums[j]], s_nxt[j, :particle_nums[j]])
        s_cur = s_pred
    loss = loss / (n_roll * B)
    loss.backward()
    out = {"nf_effect": np.asarray(NF), "states": states, "states_delta": states_delta, "attrs": attrs, "dens": dens,
           "particle_nums": particle_nums, "loss": np.asarray(loss.item()), "preds": np.stack(preds, 1),
           "g_s0": s0.grad.numpy().copy(), "g_states_delta": sd.grad.numpy().copy()}
    for k, p in model.named_parameters():
        out["w/" + k] = p.detach().numpy().copy()
        out["g/" + k] = p.grad.numpy().copy()
    return out


if __name__ == "__main__":
    main()
