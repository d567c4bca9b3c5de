"""CPU-side checks: the C ABI library loads and exports what include/pile_gnn.h declares, the drop-in
classes keep the reference's checkpoint layout, and the host-side helpers agree with the oracle."""
import itertools
import os
import re

import numpy as np
import pytest
import torch

import dyn_res_pile_manip_b200 as P
from dyn_res_pile_manip_b200 import _lib, ops, synthetic
from oracle import pile_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.isfile(_lib.LIB_PATH):
        _lib.build()
    return _lib.load()


def test_library_exports_every_declared_symbol(lib):
    header = open(os.path.join(ROOT, "include", "pile_gnn.h")).read()
    declared = set(re.findall(r"\b(pile_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 20
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.pile_abi_version() == _lib.ABI_VERSION == 4
    assert lib.pile_nf_effect() == 64 and lib.pile_max_relations() == 10


def test_argument_validation_without_gpu(lib):
    assert lib.pile_step_scratch_bytes(0, 10) == -1
    assert lib.pile_step_scratch_bytes(4, 100) > 4 * 1000 * 64 * 4
    assert lib.pile_tape_step_bytes(4, 100) > 0
    assert lib.pile_predict_step(None, None, None, None, None, None, 0.08, 4, 100, None, None, None, None) != 0
    assert lib.pile_mppi_num_chunks(1000) == 8


def test_general_width_weight_pack_layout(lib):
    """Packed buffer of the general-width engine (csrc/general.cu): every matrix zero-padded to 64 * ceil(H / 64),
    forward images transposed, backward images in checkpoint layout; offsets come from the library."""
    from oracle import pile_oracle as O
    for H in (96, 150, 64):
        Hp = (H + 63) // 64 * 64
        W = O.weights_from_seed(3, nf=H)
        pack = ops.pack_weights_general(W, torch.device("cpu"), H)
        off = {n: lib.pile_general_wpack_slot_offset(i, H) for i, n in enumerate(ops.GENERAL_SLOTS)}
        assert pack.numel() == lib.pile_general_wpack_slot_offset(len(ops.GENERAL_SLOTS), H)
        rp = W["model.relation_propagator.linear.weight"]
        blk = pack[off["W_ST"]:off["W_ST"] + Hp * Hp].view(Hp, Hp)
        assert torch.equal(blk[:H, :H], rp[:, 2 * H:3 * H].t())
        assert float(blk[H:].abs().sum()) == 0 and float(blk[:, H:].abs().sum()) == 0
        assert torch.equal(pack[off["W_S"]:off["W_S"] + Hp * Hp].view(Hp, Hp)[:H, :H], rp[:, 2 * H:3 * H])
        assert torch.equal(pack[off["WD_RP"]:off["WD_RP"] + H], rp[:, 3 * H])
        pe0 = W["model.particle_encoder.model.0.weight"]
        assert torch.equal(pack[off["W_PE0T"]:off["W_PE0T"] + 8 * Hp].view(8, Hp)[:5, :H], pe0.t())
        v1 = W["model.particle_predictor.linear_1.weight"]
        assert torch.equal(pack[off["W_V1"]:off["W_V1"] + 4 * Hp].view(4, Hp)[:3, :H], v1)
    assert lib.pile_general_wpack_slot_offset(0, 300) == -1 and lib.pile_general_tape_bytes(2, 50, 300) == -1
    assert lib.pile_general_grad_offset(18, 150) == sum(v.numel() for v in O.weights_from_seed(0, nf=150).values())
    assert lib.pile_general_forward(None, 150, None, None, None, None, None, 0.08, 2, 50, None, None, None) != 0


def test_weight_pack_layout(lib, golden_weights):
    pack = ops.pack_weights(golden_weights, torch.device("cpu"))
    assert pack.numel() == lib.pile_wpack_total()
    H = 64
    off = {n: lib.pile_wpack_slot_offset(i) for i, n in enumerate(ops.WSLOTS)}
    rp = golden_weights["model.relation_propagator.linear.weight"]
    assert torch.equal(pack[off["W_RT"]:off["W_RT"] + H * H].view(H, H), rp[:, H:2 * H].t())
    assert torch.equal(pack[off["WD_RP"]:off["WD_RP"] + H], rp[:, 3 * H])
    assert torch.equal(pack[off["W_S"]:off["W_S"] + H * H].view(H, H), rp[:, 2 * H:3 * H])
    pe0 = golden_weights["model.particle_encoder.model.0.weight"]
    assert torch.equal(pack[off["W_PE0T"]:off["W_PE0T"] + 8 * H].view(8, H)[:5], pe0.t())
    assert pack[off["W_PE0T"]:off["W_PE0T"] + 8 * H].view(8, H)[5:].abs().sum() == 0


def test_checkpoint_layout_and_init_match_reference(golden_weights):
    torch.manual_seed(0)
    m = P.PropNetDiffDenModel(synthetic.default_config(), False)
    sd = m.state_dict()
    assert list(sd.keys()) == list(golden_weights.keys())
    for k in sd:
        assert torch.equal(sd[k], golden_weights[k]), k
    m2 = P.PropNetDiffDenModel(synthetic.default_config())
    missing = m2.load_state_dict(golden_weights, strict=False)
    assert not missing.missing_keys and not missing.unexpected_keys


def test_no_cpu_fallback():
    m = P.PropNetDiffDenModel(synthetic.default_config())
    s = torch.zeros(1, 5, 3)
    with pytest.raises(_lib.PileLibraryError):
        m.predict_one_step(torch.zeros(1, 5), s, s, torch.ones(1))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "dyn_res_pile_manip_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", text, re.M), f


def test_sort10_network_is_a_sorting_network():
    src = open(os.path.join(ROOT, "dyn_res_pile_manip_b200", "csrc", "nbr.cu")).read()
    body = src[src.index("void sort10"):src.index("constexpr int NBR_THREADS")]
    net = [(int(a), int(b)) for a, b in re.findall(r"PILE_CE\((\d+), (\d+)\)", body)]
    assert len(net) == 29
    for bits in itertools.product([0, 1], repeat=10):
        a = list(bits)
        for i, j in net:
            if a[i] > a[j]:
                a[i], a[j] = a[j], a[i]
        assert a == sorted(a)


def test_dense_relation_shim_roundtrip():
    s, _ = synthetic.make_pile_batch(3, 40, seed=4)
    adj = O.adjacency(torch.from_numpy(s), torch.zeros(3, 40, 3), 0.08)
    Rr, Rs = O.one_hot_relations(adj)
    rel = ops.Relations.from_dense(Rr, Rs)
    for b, (r, c) in enumerate(O.edge_lists(adj)):
        got = rel.edge_sets()[b]
        assert np.array_equal(got[:, 0], r) and np.array_equal(got[:, 1], c)
    Rr2, Rs2 = rel.to_dense()
    assert torch.equal(Rr2, Rr) and torch.equal(Rs2, Rs)
    # shuffled relation order + padding rows still give receiver-grouped lists
    perm = torch.randperm(Rr.shape[1])
    rel2 = ops.Relations.from_dense(Rr[:, perm], Rs[:, perm])
    assert torch.equal(rel2.rowptr, rel.rowptr) and torch.equal(rel2.col, rel.col)


def test_planner_host_pieces_match_oracle(golden):
    env, cfg = synthetic.FakeEnv(), synthetic.default_config()
    planner = P.PlannerGD(cfg, env)
    np.random.seed(12)
    mine = planner.sample_action_sequences(golden["E/init"], np.zeros(5), 16, None, None)
    np.testing.assert_allclose(mine, golden["E/sampled"], rtol=0, atol=1e-12)
    assert planner.cam12 == ops.cam_matrix12(golden["cam_extrinsic"])
    np.testing.assert_allclose(np.array(planner.cam12).reshape(3, 4),
                               O.world_to_cam_matrix(golden["cam_extrinsic"]).numpy()[:3])
    assert P.particle_num_to_iter_time(100) == 72 and P.particle_num_to_iter_time(300) == 742
    planner.particle_num = 60
    coor = planner.goal_coordinates(synthetic.make_goal("bar"), "cpu")
    assert np.array_equal(coor.numpy(), golden["D/goal_coor"])


def test_synthetic_pile_statistics():
    s, d = synthetic.make_pile(100, seed=0)
    assert s.shape == (100, 3) and 800 < d < 2500
    adj = O.adjacency(torch.from_numpy(s)[None], torch.zeros(1, 100, 3), 0.08)
    deg = adj[0].sum(1)
    assert deg.max() <= 10 and deg.min() >= 1 and adj[0].diagonal().all()


def test_warm_start_shift_is_the_reference_expression():
    """flex_env.py:1112-1113: a length-1 initial sequence comes back unchanged; for longer ones the reference
    concatenates action_full[1:] ([traj-1, 4]) with a [steps, traj, 4] array, which numpy rejects - kept as is."""
    from dyn_res_pile_manip_b200.observation import shift_warm_start
    one = np.zeros((1, 6, 4))
    assert shift_warm_start(one, np.ones((6, 4), np.float32), 1) is one
    with pytest.raises(ValueError):
        shift_warm_start(np.zeros((3, 6, 4)), np.ones((6, 4), np.float32), 1)


def test_synthetic_observation_layout():
    """render_observation produces what FlexEnv.obs2ptcl_fixed_num_batch asserts on (flex_env.py:934-940)."""
    from dyn_res_pile_manip_b200 import synthetic
    env = synthetic.FakeEnv()
    st, _ = synthetic.make_pile_batch(1, 40, seed=2)
    obs = synthetic.render_observation(st[0], env)
    assert obs.shape == (env.screenHeight, env.screenWidth, 5) and obs.dtype == np.float32
    assert obs[..., :3].max() <= 255.0 and obs[..., :3].min() >= 0.0 and obs[..., :3].max() >= 1.0
    assert 0.7 * env.global_scale <= obs[..., -1].max() <= 0.8 * env.global_scale
    depth = obs[..., -1] / env.global_scale
    fg = depth < 0.599 / 0.8
    assert 0 < fg.sum() < fg.size // 4           # a pile in the middle of an empty table


def test_from_dense_reports_the_degree_and_rejects_overfull_samples():
    """Dense Rr / Rs with more than ten relations into one particle are accepted (the model routes them to the general
    kernels of the training path); more than 10 N relations per sample do not fit the fixed-capacity lists."""
    N = 14
    rr = torch.zeros(1, 12, N)
    rs = torch.zeros(1, 12, N)
    rr[0, :, 3] = 1.0                                  # twelve relations into particle 3
    rs[0, torch.arange(12), torch.arange(12)] = 1.0
    rel = ops.Relations.from_dense(rr, rs)
    assert rel.max_degree == 12 and rel.rowptr[0, 4].item() - rel.rowptr[0, 3].item() == 12
    rr[0, 10:, 3] = 0.0
    rr[0, 10:, 4] = 1.0                                # ten into particle 3, two into particle 4
    rel = ops.Relations.from_dense(rr, rs)
    assert rel.max_degree == 10 and rel.rowptr[0, 4].item() - rel.rowptr[0, 3].item() == 10
    with pytest.raises(ValueError):
        ops.Relations.from_dense(torch.zeros(1, 10 * N + 1, N), torch.zeros(1, 10 * N + 1, N))


def test_regressor_module_matches_reference_layout_and_init():
    """dyn_res_pile_manip_b200.regressor.MPCResRgrNoPool: the reference's state_dict keys and, under the same seed,
    its initial weights (oracle.regressor_oracle.build_network reproduces the reference's construction, pinned by
    tests/golden/golden_rgr_v1.npz); the input builder is the reference's cv2 sequence."""
    from oracle import regressor_oracle as RO
    from dyn_res_pile_manip_b200.regressor import MPCResRgrNoPool, regressor_input
    g = np.load(os.path.join(ROOT, "tests", "golden", "golden_rgr_v1.npz"))
    torch.manual_seed(int(g["seed"]))
    mine = MPCResRgrNoPool({"train_res_cls": {"state_h": 224, "state_w": 224, "res_dim": 6}})
    ref = RO.build_network(int(g["seed"]))
    sd, rsd = mine.state_dict(), ref.state_dict()
    assert list(sd) == ["model." + k for k in rsd]
    assert list(sd)[0] == "model.0.weight" and list(sd)[-1] == "model.19.bias"
    for k, v in rsd.items():
        assert torch.equal(sd["model." + k], v), k
    x = regressor_input(g["fg"].astype(np.float32), g["goal"].astype(np.float32), 224, 224)
    assert np.array_equal(x, g["x"])
    lib = _lib.load()
    assert lib.pile_rgr_param_offset(20) == sum(v.numel() for v in sd.values()) == 114_193_217
    assert lib.pile_rgr_param_offset(1) == 6 * 64 * 16
    assert lib.pile_rgr_workspace_bytes(1, 224, 224) > 0 and lib.pile_rgr_workspace_bytes(1, 200, 200) == -1


def test_goal_cache_is_keyed_on_content():
    """two different goals of the same shape at the same address must not share a cache entry; the same
    contents in a fresh tensor must hit it."""
    from dyn_res_pile_manip_b200.rewards import GoalCache, shape_goal_image
    cache = GoalCache()
    bar, disc = synthetic.make_goal("bar"), synthetic.make_goal("disc")
    t = torch.from_numpy(bar.copy())
    img_bar = cache.shaped(t)
    assert torch.equal(img_bar, shape_goal_image(torch.from_numpy(bar)))
    t.copy_(torch.from_numpy(disc))                          # same storage, same shape, new contents (in-place)
    img_disc = cache.shaped(t)
    assert torch.equal(img_disc, shape_goal_image(torch.from_numpy(disc))) and not torch.equal(img_disc, img_bar)
    assert cache.shaped(torch.from_numpy(disc.copy())) is img_disc          # equal contents, fresh tensor: hit
    img2 = cache.shaped_np(bar, torch.from_numpy(bar))                       # numpy entry: hash of the bytes
    assert torch.equal(img2, img_bar)
    assert cache.shaped_np(bar.copy(), torch.from_numpy(bar.copy())) is img2
    assert not torch.equal(cache.shaped_np(disc, torch.from_numpy(disc)), img2)


def test_workspace_keeps_only_recent_sizes(lib):
    """the dynamic-resolution loop changes N every step; scratch must not accumulate."""
    ws = ops.Workspace()
    dev = torch.device("cpu")
    bufs = [ws.scratch(2, n, dev) for n in (10, 11, 12, 13, 14)]
    assert len(ws._scratch) == ops.Workspace.KEEP
    assert ws.scratch(2, 14, dev) is bufs[-1] and ws.scratch(2, 13, dev) is bufs[-2]
    assert ws.scratch(2, 10, dev) is not bufs[0]            # evicted and re-created
    for n in (10, 11, 12, 13):
        ws.bwd(2, n, dev)
    assert len(ws._bwd) == ops.Workspace.KEEP
    with pytest.raises(_lib.PileLibraryError):
        ws.scratch(0, 5, dev)
    with pytest.raises(_lib.PileLibraryError):
        ws.bwd(3, -1, dev)


def test_pusher_struct_layout_and_planner_frames():
    """`pile_pusher` as ctypes sees it (kind, 12 matrix floats, global_scale, s2r_scale, centre) and the two planner
    frames: simulator camera and real robot (planners.py:192-257 / 259-300, dispatch :345-348)."""
    import ctypes
    assert ctypes.sizeof(_lib.PusherStruct) == 4 + 12 * 4 + 4 * 4
    env = synthetic.FakeEnv()
    sim = P.PlannerGD(synthetic.default_config(), env)
    assert sim.pusher.kind == 0 and sim.pusher.struct.global_scale == synthetic.GLOBAL_SCALE
    assert [sim.pusher.struct.cam_m12[i] for i in range(12)] == sim.cam12 and sim.reward_offset() == (0., 0.)

    class Real(synthetic.FakeEnv):
        is_real = True
        s2r_scale, wkspc_center_x, wkspc_center_y = 10.0, 0.03, -0.02
        crop_w_lower, crop_w_off, crop_h_lower, crop_h_off = 100, 20, 80, 10
    real = P.PlannerGD(synthetic.default_config(), Real())
    assert real.pusher.kind == 1 and real.pusher.struct.s2r_scale == 10.0
    assert abs(real.pusher.struct.wkspc_center_x - 0.03) < 1e-7 and real.reward_offset() == (-80.0, -70.0)
    assert real.pusher.signature() != sim.pusher.signature()
