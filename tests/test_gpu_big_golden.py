"""Parity at BASELINE.json's sizes against vectors the REAL reference produced (tests/golden/make_golden_big.py):
one step at 8 x 300, full-horizon rollouts at 100 x T=10 and 300 x T=20 with the drift curve (per-step
||d|| / ||s||), the relation-set Jaccard index per step, rewards and action gradients -- on all three GEMM engines.

Cases G/H run the random-init network as it is: it translates the whole pile by ~0.25 per step (out of the camera's
view after a few steps, through z = 0 around step 8), so they pin the arithmetic far away from a pile at rest.
Cases J/K scale the predictor's output layer by 0.02: the pile stays in place, pushes matter and the action
gradient is dense -- the regime of a trained checkpoint."""
import numpy as np
import pytest
import torch

import dyn_res_pile_manip_b200 as P
from dyn_res_pile_manip_b200 import ops, synthetic
from conftest import tamed_weights

pytestmark = pytest.mark.gpu
DEV = "cuda"
POS_TOL = 1e-4


@pytest.fixture(autouse=True, params=["fp32", "tc", "tc_tmem"])
def engine(request):
    old = ops.set_tensor_cores({"fp32": 0, "tc": 1, "tc_tmem": 2}[request.param])
    yield request.param
    ops.set_tensor_cores(old)


def cuda(x):
    return torch.as_tensor(x).to(DEV)


def relerr(a, b):
    a, b = torch.as_tensor(a).detach().double().cpu(), torch.as_tensor(b).detach().double().cpu()
    return float((a - b).norm() / b.norm())


def model_with(weights):
    m = P.PropNetDiffDenModel(synthetic.default_config(), True)
    m.load_state_dict(weights)
    return m.to(DEV).requires_grad_(False)


def pair_set(coo):
    c = np.asarray(coo).astype(np.int64)
    return set((c[:, 0] << 40 | c[:, 1] << 20 | c[:, 2]).tolist())


def coo_from_relations(rel):
    rows = []
    for b, e in enumerate(rel.edge_sets()):
        rows.append(np.concatenate([np.full((e.shape[0], 1), b), e], axis=1))
    return np.concatenate(rows).astype(np.int16)


def drift_report(planner, model, g, tag):
    """Free-running rollout of case `tag` -> (per-step relative position error, per-step Jaccard index of the
    relation sets against the reference's, states, action tensor with grad)."""
    nb, ns, N, T = [int(v) for v in g[tag + "/dims"]]
    planner.particle_num = N
    acts = cuda(g[tag + "/acts"]).requires_grad_(True)
    out = planner.ptcl_model_rollout(cuda(g[tag + "/s0"]), cuda(g[tag + "/dens"]), torch.zeros(nb, N, device=DEV), model, acts)
    pred = out["model_rollout"]["state_pred"]
    drift = [relerr(pred[:, t], g[tag + "/state_pred"][:, t]) for t in range(T)]
    jac = []
    s = cuda(g[tag + "/s0"]).repeat(ns, 1, 1)
    for t in range(T):
        with torch.no_grad():
            sd = planner.gen_s_delta(s, acts.detach()[:, t].contiguous())
            mine = pair_set(coo_from_relations(ops.build_relations(s, sd, 0.08)))
        ref = pair_set(g[tag + "/rel%d" % t])
        jac.append(len(mine & ref) / len(mine | ref))
        s = pred.detach()[:, t].contiguous()
    return drift, jac, pred, acts


def test_one_step_at_300_particles(golden_big, golden_weights, engine):
    g = golden_big
    model = model_with(golden_weights)
    planner = P.PlannerGD(synthetic.default_config(), synthetic.FakeEnv())
    planner.particle_num = 300
    sd = planner.gen_s_delta(cuda(g["F/s_cur"]), cuda(g["F/act"]))
    np.testing.assert_allclose(sd.cpu().numpy(), g["F/s_delta"], rtol=0, atol=2e-7)
    out = model.predict_one_step(torch.zeros(8, 300, device=DEV), cuda(g["F/s_cur"]), cuda(g["F/s_delta"]), cuda(g["F/dens"]))
    assert np.array_equal(coo_from_relations(model.relations_of_last_step()), g["F/rel"])
    err = relerr(out, g["F/s_pred"])
    print("one step 8x300 [%s]: rel err %.2e, max abs %.2e" % (engine, err, float((out.cpu() - torch.from_numpy(g["F/s_pred"])).abs().max())))
    assert err < POS_TOL
    np.testing.assert_allclose(out.cpu().numpy(), g["F/s_pred"], rtol=0, atol=5e-6 if engine == "fp32" else 6e-5)


@pytest.mark.parametrize("tag", ["J", "K", "G", "H"])
def test_full_horizon_drift_and_gradients(golden_big, golden_weights, engine, tag):
    g = golden_big
    nb, ns, N, T = [int(v) for v in g[tag + "/dims"]]
    tame = float(g[tag + "/tame"])
    model = model_with(tamed_weights(golden_weights, tame))
    planner = P.PlannerGD(synthetic.default_config(), synthetic.FakeEnv())
    drift, jac, pred, acts = drift_report(planner, model, g, tag)
    print("case %s (%d x %d particles x T=%d, tame=%g) [%s]" % (tag, nb * ns, N, T, tame, engine))
    print("  drift   " + " ".join("%.1e" % d for d in drift))
    print("  jaccard " + " ".join("%.4f" % j for j in jac))
    assert drift[0] < POS_TOL and jac[0] == 1.0             # the north-star bar: one step
    if tame < 1.0:
        # pile at rest: the whole horizon stays inside the one-step bar, relation sets identical or one flip away
        assert max(drift) < POS_TOL, drift
        assert min(jac) > 0.999, jac
    else:
        # the random-init network is a chaotic map far from a pile at rest: drift is reported, bounded loosely
        assert max(drift) < 2e-2, drift
    goal = cuda(synthetic.make_goal(str(g[tag + "/goal_kind"])))
    obs = pred.reshape(ns * nb, 1, T, N, 3).permute(0, 2, 1, 3, 4)
    reward, next_r = planner.ptcl_evaluate_traj(obs, goal, cuda(g[tag + "/goal_coor"]))
    if tame < 1.0:
        np.testing.assert_allclose(next_r.detach().cpu().numpy(), g[tag + "/next_r"], rtol=2e-5 if engine == "fp32" else 3e-4)
        torch.sum(-reward).backward()
        got, ref = acts.grad.cpu().numpy(), g[tag + "/act_grad"]
        gerr = float(np.linalg.norm(got - ref) / np.linalg.norm(ref))
        print("  action gradient rel err %.2e (max |ref| %.3g)" % (gerr, np.abs(ref).max()))
        # fp32 engine: summation order only; tensor engine: sign bits of ReLUs recomputed from 2^-16-rounded operands
        assert gerr < (2e-3 if engine == "fp32" else 2e-2), gerr


def test_relation_sets_with_reference_states_reinjected(golden_big, golden_weights):
    """Teacher-forced: the relation search driven by the reference's own per-step states must give the reference's
    relation sets at every step of the T=20 rollout at 300 particles (no tolerance)."""
    g = golden_big
    planner = P.PlannerGD(synthetic.default_config(), synthetic.FakeEnv())
    for tag in ("H", "K"):
        nb, ns, N, T = [int(v) for v in g[tag + "/dims"]]
        planner.particle_num = N
        s = cuda(g[tag + "/s0"]).repeat(ns, 1, 1)
        for t in range(T):
            sd = planner.gen_s_delta(s, cuda(g[tag + "/acts"][:, t]))
            rel = ops.build_relations(s, sd, 0.08)
            assert np.array_equal(coo_from_relations(rel), g[tag + "/rel%d" % t]), (tag, t)
            s = cuda(g[tag + "/state_pred"][:, t])
