"""Parity of the CUDA path (through the drop-in API -> C ABI) against the golden vectors produced by the
reference and against the CPU oracle.  Relation sets: bit-exact.  Positions: ||d|| / ||s|| <= 1e-4 after one
step (BASELINE.json north_star); everything float is far inside that here because the whole path is fp32."""
import numpy as np
import pytest
import torch

import dyn_res_pile_manip_b200 as P
from dyn_res_pile_manip_b200 import _lib, ops, synthetic
from oracle import pile_oracle as O

pytestmark = pytest.mark.gpu
POS_TOL = 1e-4       # north_star: positions within 1e-4 relative after one step
DEV = "cuda"


@pytest.fixture(autouse=True, params=["fp32", "tc", "tc_tmem"])
def engine(request):
    """Every parity test runs on all GEMM engines: FP32 CUDA-core tiles, tcgen05 (bf16 hi/lo split) tiles with
    shared-memory activations, and the same with the relation encoder's activations in tensor memory."""
    old = ops.set_tensor_cores({"fp32": 0, "tc": 1, "tc_tmem": 2}[request.param])
    yield request.param
    ops.set_tensor_cores(old)


def atol_for(engine):
    # fp32 engine: summation-order noise only; tensor-core engine: 3-pass split keeps ~2^-16 per product
    return 5e-6 if engine == "fp32" else 4e-5


def relerr(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).norm() / b.norm())


def coo_from_relations(rel):
    rows = []
    for b, e in enumerate(rel.edge_sets()):
        rows.append(np.concatenate([np.full((e.shape[0], 1), b), e], axis=1))
    return np.concatenate(rows).astype(np.int16)


@pytest.fixture(scope="module")
def model(golden_weights):
    m = P.PropNetDiffDenModel(synthetic.default_config(), True)
    m.load_state_dict(golden_weights)
    # frozen parameters: these tests exercise the planner-side engines (direct predict_one_step calls with trainable
    # parameters under autograd take the training kernels instead, tests/test_gpu_training.py)
    return m.to(DEV).requires_grad_(False)


@pytest.fixture(scope="module")
def planner():
    return P.PlannerGD(synthetic.default_config(), synthetic.FakeEnv())


def cuda(x):
    return torch.as_tensor(x).to(DEV)


def test_s_delta_matches_reference(golden, planner):
    planner.particle_num = 100
    sd = planner.gen_s_delta(cuda(golden["A/s_cur"]), cuda(golden["A/act"]))
    np.testing.assert_allclose(sd.cpu().numpy(), golden["A/s_delta"], rtol=0, atol=2e-7)


@pytest.mark.parametrize("case", ["A", "B", "C"])
def test_relation_sets_bit_exact_vs_reference(golden, case):
    nums = golden.get(case + "/nums")
    rel = ops.build_relations(cuda(golden[case + "/s_cur"]), cuda(golden[case + "/s_delta"]), 0.08, nums)
    assert np.array_equal(coo_from_relations(rel), golden[case + "/rel"])


@pytest.mark.parametrize("N,B", [(1, 2), (9, 3), (10, 3), (11, 3), (50, 8), (100, 16), (300, 8), (777, 2)])
def test_relation_sets_bit_exact_vs_oracle(N, B):
    rng = np.random.RandomState(N)
    s, _ = synthetic.make_pile_batch(B, N, seed=N) if N >= 10 else (rng.uniform(-.1, .1, (B, N, 3)).astype(np.float32), 0)
    sd = (rng.normal(0, 0.02, size=s.shape) * (rng.uniform(size=s.shape[:2] + (1,)) < 0.3)).astype(np.float32)
    adj = O.adjacency(torch.from_numpy(s), torch.from_numpy(sd), 0.08)
    rel = ops.build_relations(cuda(s), cuda(sd), 0.08)
    assert np.array_equal(coo_from_relations(rel), adj.nonzero().to(torch.int16).numpy())
    deg = (rel.rowptr[:, 1:] - rel.rowptr[:, :-1]).cpu()
    assert deg.min() >= 1 and deg.max() <= min(10, N)


@pytest.mark.parametrize("N,B", [(1, 2), (7, 3), (10, 3), (31, 4), (32, 4), (33, 5), (64, 3), (100, 4), (300, 3), (341, 2), (500, 2)])
def test_relation_search_block_borders_and_padding_mask(N, B):
    """The search takes candidates 32 at a time (admission mask) with the SoA copy padded to a multiple of 32: sizes
    around the block borders, with the particle_nums padding mask, against the oracle."""
    rng = np.random.RandomState(100 + N)
    s = rng.uniform(-.12, .12, (B, N, 3)).astype(np.float32)
    s[..., 2] += 0.74
    sd = (rng.normal(0, 0.02, size=s.shape) * (rng.uniform(size=s.shape[:2] + (1,)) < 0.3)).astype(np.float32)
    nums = np.array([N] + [max(1, N - 3 - b) for b in range(1, B)], dtype=np.int32)
    adj = O.adjacency(torch.from_numpy(s), torch.from_numpy(sd), 0.08, nums)
    rel = ops.build_relations(cuda(s), cuda(sd), 0.08, nums)
    assert np.array_equal(coo_from_relations(rel), adj.nonzero().to(torch.int16).numpy())


def test_largest_sample_and_one_past_it():
    """The relation search keeps a whole sample in 200 KB of shared memory (184 bytes per particle): 1112 particles is
    the largest sample it takes -- checked against the oracle, with and without the transposed lists -- and 1113 is
    refused loudly instead of launching with too little shared memory."""
    N = 1112
    rng = np.random.RandomState(3)
    s = rng.uniform(-.25, .25, (2, N, 3)).astype(np.float32)
    s[..., 2] = 0.74 + 0.02 * rng.uniform(size=(2, N)).astype(np.float32)
    sd = (rng.normal(0, 0.02, size=s.shape) * (rng.uniform(size=s.shape[:2] + (1,)) < 0.3)).astype(np.float32)
    adj = O.adjacency(torch.from_numpy(s), torch.from_numpy(sd), 0.08)
    want = adj.nonzero().to(torch.int16).numpy()
    rel = ops.build_relations(cuda(s), cuda(sd), 0.08)
    assert np.array_equal(coo_from_relations(rel), want)
    rel_t, _ = ops.build_relations(cuda(s), cuda(sd), 0.08, with_transpose=True)
    assert np.array_equal(coo_from_relations(rel_t), want)
    big = np.zeros((1, N + 1, 3), np.float32)
    with pytest.raises(_lib.PileLibraryError):
        ops.build_relations(cuda(big), cuda(big), 0.08)


@pytest.mark.parametrize("N,B", [(7, 2), (40, 3), (100, 4), (300, 3), (341, 2)])
def test_sender_major_transpose(N, B):
    """The transposed relation lists the backward gathers over: for every sender its receivers in ascending order and
    the ids of those relations (built by a counting sort + per-segment sort in shared memory)."""
    rng = np.random.RandomState(7 + N)
    s = rng.uniform(-.12, .12, (B, N, 3)).astype(np.float32)
    if N >= 40:
        s[:, 3:20, :2] = s[:, 0:1, :2] + rng.normal(0, 1e-3, (B, 17, 2)).astype(np.float32)     # a hub: large in-degrees
    nums = np.array([N] + [max(1, N - 2 - b) for b in range(1, B)], dtype=np.int32)
    rel, (trp, trecv, tedge) = ops.build_relations(cuda(s), cuda(np.zeros_like(s)), 0.08, nums, with_transpose=True)
    rp, col, row = rel.rowptr.cpu().numpy(), rel.col.cpu().numpy(), rel.row.cpu().numpy()
    trp, trecv, tedge = trp.cpu().numpy(), trecv.cpu().numpy(), tedge.cpu().numpy()
    for b in range(B):
        ne = rp[b, -1]
        assert trp[b, -1] == ne and trp[b, 0] == 0
        order = np.lexsort((row[b, :ne], col[b, :ne]))          # by sender, then receiver
        assert np.array_equal(tedge[b, :ne], order)
        assert np.array_equal(trecv[b, :ne], row[b, :ne][order])
        assert np.array_equal(np.diff(trp[b]), np.bincount(col[b, :ne], minlength=N))


@pytest.mark.parametrize("garbage", ["valid_lists", "duplicates", "out_of_range", "negative_offsets", "stale_step"])
def test_relation_search_warm_start_never_changes_the_result(garbage):
    """The search takes its first admission bound from whatever the output lists hold on entry (the relations of the
    rollout step / planner iteration before).  Any ten distinct in-range candidates give a valid bound; anything else must
    be ignored: the result is the cold-start result (empty lists on entry) whatever the buffers contained -- including
    sample 0, where 15 coincident points tie at the tenth place of many receivers (lower index wins; torch.topk's own
    choice among such ties is unspecified, so only the tie-free samples are also compared with the oracle)."""
    from dyn_res_pile_manip_b200 import _lib
    B, N = 3, 77
    rng = np.random.RandomState(5)
    s = rng.uniform(-.12, .12, (B, N, 3)).astype(np.float32)
    s[0, 5:19] = s[0, 4]                                     # coincident points: ties at the tenth place
    want = O.adjacency(torch.tensor(s), torch.zeros(B, N, 3), 0.08)
    rowptr = torch.arange(N + 1, dtype=torch.int32, device=DEV).repeat(B, 1) * 10
    col = torch.zeros(B, 10 * N, dtype=torch.int32, device=DEV)
    if garbage == "valid_lists":          # ten distinct random candidates per receiver, ascending: a (loose) valid bound
        col = torch.tensor(np.stack([np.concatenate([np.sort(rng.choice(N, 10, replace=False)) for _ in range(N)])
                                     for _ in range(B)]).astype(np.int32), device=DEV)
    elif garbage == "duplicates":         # in range but not distinct: would give a bound that is too tight
        col = torch.tensor(rng.randint(0, 3, (B, 10 * N)).astype(np.int32), device=DEV)
    elif garbage == "out_of_range":
        col = torch.tensor(rng.randint(-5 * N, 5 * N, (B, 10 * N)).astype(np.int32), device=DEV)
    elif garbage == "negative_offsets":
        rowptr = torch.tensor(rng.randint(-2 ** 31, 2 ** 31 - 1, (B, N + 1), dtype=np.int64).astype(np.int32), device=DEV)
    elif garbage == "stale_step":         # the lists of a different configuration of the same pile
        other = ops.build_relations(cuda(s[:, ::-1].copy()), cuda(np.zeros_like(s)), 0.08)
        rowptr, col = other.rowptr.clone(), other.col.clone()
    row = torch.zeros(B, 10 * N, dtype=torch.int32, device=DEV)
    sc, sd = cuda(s), cuda(np.zeros_like(s))
    _lib.check(_lib.load().pile_build_relations(_lib.ptr(sc), _lib.ptr(sd), None, B, N, 0.08, _lib.ptr(rowptr), _lib.ptr(col),
                                                _lib.ptr(row), None, None, None, ops._stream()), "pile_build_relations")
    got = ops.Relations(rowptr, col, row).edge_sets()
    cold_rp = torch.zeros(B, N + 1, dtype=torch.int32, device=DEV)          # every old list empty: no warm start
    cold_col, cold_row = torch.zeros_like(col), torch.zeros_like(row)
    _lib.check(_lib.load().pile_build_relations(_lib.ptr(sc), _lib.ptr(sd), None, B, N, 0.08, _lib.ptr(cold_rp),
                                                _lib.ptr(cold_col), _lib.ptr(cold_row), None, None, None, ops._stream()),
               "pile_build_relations")
    cold = ops.Relations(cold_rp, cold_col, cold_row).edge_sets()
    for b in range(B):
        assert np.array_equal(got[b], cold[b]), (garbage, b)
        if b > 0:
            assert np.array_equal(got[b], np.argwhere(want[b].numpy() > 0)), (garbage, b)
    inside = got[0][got[0][:, 0] == 11, 1]                   # a receiver inside the coincident group keeps the lowest ten
    assert list(inside) == list(range(4, 14))


def test_relations_duplicate_points_lowest_index_wins():
    # 14 coincident particles: every distance ties at 0 -> the 10 lowest sender indices must be kept
    s = np.zeros((1, 14, 3), dtype=np.float32)
    rel = ops.build_relations(cuda(s), cuda(np.zeros_like(s)), 0.08)
    e = rel.edge_sets()[0]
    for i in range(14):
        assert list(e[e[:, 0] == i, 1]) == list(range(10))


@pytest.mark.parametrize("case", ["A", "B", "C"])
def test_one_step_positions_vs_reference(golden, model, case, engine):
    g = {k.split("/")[1]: v for k, v in golden.items() if k.startswith(case + "/")}
    B, N, _ = g["s_cur"].shape
    a = cuda(g["a_cur"]) if "a_cur" in g else torch.zeros(B, N, device=DEV)
    out = model.predict_one_step(a, cuda(g["s_cur"]), cuda(g["s_delta"]), cuda(g["dens"]), g.get("nums"))
    assert relerr(out, g["s_pred"]) < POS_TOL
    valid = slice(None)
    np.testing.assert_allclose(out.cpu().numpy()[valid], g["s_pred"][valid], rtol=0, atol=atol_for(engine))
    assert np.array_equal(coo_from_relations(model.relations_of_last_step()), g["rel"])


def test_forward_accepts_dense_one_hot_relations(golden, model, golden_weights, engine):
    s, sd, dn = (torch.from_numpy(golden["A/" + k]) for k in ("s_cur", "s_delta", "dens"))
    adj = O.adjacency(s, sd, 0.08)
    Rr, Rs = O.one_hot_relations(adj)
    perm = torch.randperm(Rr.shape[1])
    out = model.model.forward(torch.zeros(4, 100, device=DEV), s.to(DEV), sd.to(DEV), Rr[:, perm].to(DEV),
                              Rs[:, perm].to(DEV), dn.to(DEV))
    np.testing.assert_allclose(out.cpu().numpy(), golden["A/s_pred"], rtol=0, atol=atol_for(engine))
    rel = ops.build_relations(s.to(DEV), sd.to(DEV), 0.08)
    out2 = model.model.forward(torch.zeros(4, 100, device=DEV), s.to(DEV), sd.to(DEV), rel, None, dn.to(DEV))
    assert torch.equal(out, out2)


def test_rollout_reward_gradient_vs_reference(golden, model, planner, engine):
    n_batch, n_sample, N, T = 2, 3, 60, 4
    planner.particle_num = N
    acts = cuda(golden["D/acts"]).requires_grad_(True)
    out = planner.ptcl_model_rollout(cuda(golden["D/s0"]), cuda(golden["D/dens"]), torch.zeros(n_batch, N, device=DEV),
                                     model, acts)
    pred = out["model_rollout"]["state_pred"]
    assert float(out["rollout_time"]) > 0
    drift = [relerr(pred[:, t], golden["D/state_pred"][:, t]) for t in range(T)]
    assert drift[0] < POS_TOL and max(drift) < 1e-4, drift
    goal = cuda(synthetic.make_goal(str(golden["D/goal_kind"])))
    obs = pred.reshape(n_sample * n_batch, 1, T, N, 3).permute(0, 2, 1, 3, 4)
    reward, next_r = planner.ptcl_evaluate_traj(obs, goal, cuda(golden["D/goal_coor"]))
    rtol = 2e-5 if engine == "fp32" else 2e-4
    np.testing.assert_allclose(reward.detach().cpu().numpy(), golden["D/reward"], rtol=rtol)
    np.testing.assert_allclose(next_r.detach().cpu().numpy(), golden["D/next_r"], rtol=rtol)
    torch.sum(-reward).backward()
    g, ref = acts.grad.cpu().numpy(), golden["D/act_grad"]
    np.testing.assert_allclose(g, ref, rtol=5e-3, atol=(5e-4 if engine == "fp32" else 3e-3) * np.abs(ref).max())


def test_rollout_relation_sets_every_step(golden, model, planner):
    """Drive the CUDA relation search with the reference's own per-step states: sets must be identical."""
    N, T = 60, 4
    planner.particle_num = N
    s = torch.from_numpy(golden["D/s0"]).repeat(3, 1, 1)
    for t in range(T):
        sd = planner.gen_s_delta(s.to(DEV), cuda(golden["D/acts"][:, t]))
        rel = ops.build_relations(s.to(DEV), sd, 0.08)
        assert np.array_equal(coo_from_relations(rel), golden["D/rel%d" % t]), t
        s = torch.from_numpy(golden["D/state_pred"][:, t])


def test_step_gradients_vs_oracle_autograd(golden, model, golden_weights):
    s = torch.tensor(golden["A/s_cur"], requires_grad=True)
    sd = torch.tensor(golden["A/s_delta"], requires_grad=True)
    rng = np.random.RandomState(0)
    a = torch.tensor(rng.uniform(0, 1, (4, 100)).astype(np.float32))
    dn = torch.from_numpy(golden["A/dens"])
    wgt = torch.tensor(rng.normal(size=(4, 100, 3)).astype(np.float32))
    (O.predict_one_step(golden_weights, 0.08, a, s, sd, dn) * wgt).sum().backward()
    s2 = s.detach().to(DEV).requires_grad_(True)
    sd2 = sd.detach().to(DEV).requires_grad_(True)
    (model.predict_one_step(a.to(DEV), s2, sd2, dn.to(DEV)) * wgt.to(DEV)).sum().backward()     # sign-bit dgrad kernels
    tol = 1e-4 if ops._lib.load().pile_get_tensor_cores() == 0 else 2e-3
    assert relerr(s2.grad, s.grad) < tol and relerr(sd2.grad, sd.grad) < tol


def test_s_delta_gradients_vs_oracle_autograd(golden, planner):
    env = synthetic.FakeEnv()
    s = torch.tensor(golden["A/s_cur"], requires_grad=True)
    act = torch.tensor(golden["A/act"], requires_grad=True)
    wgt = torch.tensor(np.random.RandomState(1).normal(size=(4, 100, 3)).astype(np.float32))
    (O.gen_s_delta(env.get_cam_extrinsics(), synthetic.GLOBAL_SCALE, s, act) * wgt).sum().backward()
    planner.particle_num = 100
    s2 = s.detach().to(DEV).requires_grad_(True)
    a2 = act.detach().to(DEV).requires_grad_(True)
    (planner.gen_s_delta(s2, a2) * wgt.to(DEV)).sum().backward()
    assert relerr(s2.grad, s.grad) < 1e-4 and relerr(a2.grad, act.grad) < 1e-4


def test_reward_vs_oracle_including_out_of_view_particles():
    rng = np.random.RandomState(3)
    goal = synthetic.make_goal("tee")
    env = synthetic.FakeEnv()
    coords = np.argwhere(goal < 0.5)[:, ::-1].astype(np.float32)
    coor, _ = synthetic.fps_np(coords, 200, 0)
    st = np.concatenate([rng.uniform(-0.4, 0.4, (6, 80, 2)), rng.uniform(0.6, 0.8, (6, 80, 1))], 2).astype(np.float32)
    s_cpu = torch.tensor(st, requires_grad=True)
    r_ref = O.reward_ptcl(s_cpu, torch.from_numpy(goal), env.get_cam_params(), torch.from_numpy(coor))
    wgt = torch.tensor(rng.normal(size=6).astype(np.float32))
    (r_ref * wgt).sum().backward()
    s_gpu = torch.tensor(st, device=DEV, requires_grad=True)
    r = P.config_reward_ptcl(s_gpu, cuda(goal), env.get_cam_params(), cuda(coor))
    np.testing.assert_allclose(r.detach().cpu().numpy(), r_ref.detach().numpy(), rtol=2e-5)
    (r * wgt.to(DEV)).sum().backward()
    assert relerr(s_gpu.grad, s_cpu.grad) < 1e-4


def test_mppi_weighting_vs_reference(golden, planner):
    out = planner.optimize_action(golden["E/sampled"], golden["E/reward"])
    np.testing.assert_allclose(out, golden["E/optimized"], rtol=2e-5, atol=1e-6)
    # many chunks + extreme reward spread (log-sum-exp path)
    rng = np.random.RandomState(2)
    acts = rng.uniform(-4, 4, (1000, 7, 1, 4))
    rew = rng.uniform(-900, -1, (1000, 1))
    np.testing.assert_allclose(planner.optimize_action(acts, rew), O.mppi_optimize_action(acts, rew, 0.1),
                               rtol=1e-4, atol=1e-5)


class _RealEnv(synthetic.FakeEnv):
    """FakeEnv with the attributes the real-robot branch reads (planners.py:271-278, 409-410)."""
    is_real = True
    crop_w_lower, crop_w_off, crop_h_lower, crop_h_off = 100, 20, 80, 10

    def __init__(self, g):
        super().__init__()
        self.wkspc_center_x, self.wkspc_center_y = float(g["wkspc_center_x"]), float(g["wkspc_center_y"])
        self.s2r_scale = float(g["s2r_scale"])


def test_gen_s_delta_irl_vs_reference():
    """Real-robot pusher model on the device (planners.py:259-300) against the reference's own output and autograd
    gradients (tests/golden/make_golden_irl.py), through a planner built for env.is_real."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_irl_v1.npz"))
    planner = P.PlannerGD(synthetic.default_config(), _RealEnv(g))
    planner.particle_num = g["s_cur"].shape[1]
    s = cuda(g["s_cur"]).requires_grad_(True)
    a = cuda(g["action"]).requires_grad_(True)
    out = planner.gen_s_delta_irl(s, a)
    np.testing.assert_allclose(out.detach().cpu().numpy(), g["s_delta"], rtol=0, atol=2e-7)
    assert (np.abs(g["s_delta"]).sum(-1) > 0).sum() > 10
    (out * cuda(g["weight"])).sum().backward()
    np.testing.assert_allclose(s.grad.cpu().numpy(), g["g_s_cur"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(a.grad.cpu().numpy(), g["g_action"], rtol=1e-4, atol=1e-5)
    assert planner.reward_offset() == (-80.0, -70.0)
    # a planner built for the simulator exposes the same method (it reads the three attributes from the env)
    sim = P.PlannerGD(synthetic.default_config(), synthetic.FakeEnv())
    sim.env.wkspc_center_x, sim.env.wkspc_center_y, sim.env.s2r_scale = planner.env.wkspc_center_x, planner.env.wkspc_center_y, planner.env.s2r_scale
    sim.particle_num = planner.particle_num
    assert torch.equal(sim.gen_s_delta_irl(s.detach(), a.detach()), out.detach())


def test_real_robot_rollout_uses_the_irl_pusher(model, golden_weights):
    """planners.py:345-348: with env.is_real the horizon rollout takes its s_delta from gen_s_delta_irl; the fused
    rollout must equal stepping the model by hand with the stand-alone irl kernel, and match the oracle."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_irl_v1.npz"))
    env = _RealEnv(g)
    planner = P.PlannerGD(synthetic.default_config(), env)
    N = g["s_cur"].shape[1]
    planner.particle_num = N
    s0, T = cuda(g["s_cur"][:2]), 3
    acts = cuda(np.stack([g["action"][[0, 1]], g["action"][[2, 3]], g["action"][[4, 0]]], axis=1))      # [2, T, 4]
    dens = torch.full((2,), 1500.0, device=DEV)
    with torch.no_grad():
        out = planner.ptcl_model_rollout(s0, dens, torch.zeros(2, N, device=DEV), model, acts)["model_rollout"]["state_pred"]
        s = s0
        for t in range(T):
            sd = planner.gen_s_delta_irl(s, acts[:, t].contiguous())
            s = model.predict_one_step(torch.zeros(2, N, device=DEV), s, sd, dens)
            assert torch.equal(s, out[:, t]), t
    s = torch.from_numpy(g["s_cur"][:2])
    for t in range(T):
        sd = O.gen_s_delta_irl(s, acts[:, t].cpu(), env.wkspc_center_x, env.wkspc_center_y, env.s2r_scale)
        s = O.predict_one_step(golden_weights, 0.08, torch.zeros(2, N), s, sd, dens.cpu())
    assert relerr(out[:, -1], s) < 1e-4
