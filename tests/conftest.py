import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "golden_v1.npz"), allow_pickle=False))


@pytest.fixture(scope="session")
def golden_weights(golden):
    import torch
    return {k[2:]: torch.from_numpy(v) for k, v in golden.items() if k.startswith("w/")}


@pytest.fixture(scope="session")
def golden_big():
    """Reference-generated vectors at BASELINE.json's particle counts (tests/golden/make_golden_big.py)."""
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "golden_big_v1.npz"), allow_pickle=False))


def tamed_weights(weights, tame):
    """Seed-0 weights with the predictor's output layer scaled (cases J/K of golden_big_v1.npz)."""
    out = {k: v.clone() for k, v in weights.items()}
    for k in ("model.particle_predictor.linear_1.weight", "model.particle_predictor.linear_1.bias"):
        out[k] = out[k] * float(tame)
    return out
