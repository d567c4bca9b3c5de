"""tcgen05 relation-encoder tiles against the FP32 CUDA-core tiles of the same library (same inputs, same
relation lists): the 3-pass bf16 hi/lo split must stay within 2^-16-level error of the fp32 result, sample by
sample, including ragged last tiles, tiny samples and the recorded ReLU sign bits."""
import numpy as np
import pytest
import torch

import dyn_res_pile_manip_b200 as P
from dyn_res_pile_manip_b200 import ops, synthetic

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def model():
    torch.manual_seed(0)
    return P.PropNetDiffDenModel(synthetic.default_config(), True).to(DEV).requires_grad_(False)     # planner-side engines


@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("N,B", [(1, 4), (7, 5), (50, 9), (100, 64), (300, 40), (13, 700), (1000, 3)])
def test_tc_step_matches_fp32_step(model, N, B, mode):
    rng = np.random.RandomState(N)
    if N >= 10:
        s, dn = synthetic.make_pile_batch(B, N, seed=N)
    else:
        s, dn = rng.uniform(-.05, .05, (B, N, 3)).astype(np.float32), np.full(B, 900, np.float32)
    sd = (rng.normal(0, 0.02, size=s.shape) * (rng.uniform(size=s.shape[:2] + (1,)) < 0.3)).astype(np.float32)
    a = rng.uniform(0, 1, (B, N)).astype(np.float32)
    args = [torch.tensor(x, device=DEV) for x in (a, s, sd, dn)]
    old = ops.set_tensor_cores(False)
    try:
        ref = model.predict_one_step(*args)
        ops.set_tensor_cores(mode)
        out = model.predict_one_step(*args)
    finally:
        ops.set_tensor_cores(old)
    assert torch.isfinite(out).all()
    disp = (ref - args[1]).abs().max().item()
    err = (out - ref).abs().max().item()
    assert err <= 2e-4 * max(disp, 1e-3), (err, disp)          # error relative to the predicted displacement
    assert float((out - ref).norm() / ref.norm()) < 2e-5        # relative to the positions (bar: 1e-4)


def test_tc_gradients_match_fp32_gradients(model):
    s, dn = synthetic.make_pile_batch(6, 120, seed=3)
    rng = np.random.RandomState(0)
    sd = rng.normal(0, 0.01, size=s.shape).astype(np.float32)
    wgt = torch.tensor(rng.normal(size=s.shape).astype(np.float32), device=DEV)
    grads = []
    for flag in (0, 2):
        old = ops.set_tensor_cores(flag)
        try:
            s_t = torch.tensor(s, device=DEV, requires_grad=True)
            sd_t = torch.tensor(sd, device=DEV, requires_grad=True)
            (model.predict_one_step(torch.zeros(6, 120, device=DEV), s_t, sd_t, torch.tensor(dn, device=DEV)) * wgt).sum().backward()
            grads.append((s_t.grad.clone(), sd_t.grad.clone()))
        finally:
            ops.set_tensor_cores(old)
    for a, b in zip(grads[0], grads[1]):
        assert float((a - b).norm() / a.norm()) < 1e-2     # ReLU sign bits may flip where a pre-activation is ~0
