"""world_size-2 `gloo` test of the sample-sharded MPPI planner's host logic.

The CUDA ops are replaced by oracle-backed test doubles (this is the one place that is legitimate: the
product itself has no CPU path); what is exercised is the real planner code: identical noise on every
rank, contiguous sample shards, one all_gather of the (2+4T) record per iteration, log-sum-exp merge --
and the requirement that 1 rank and 2 ranks produce the same plan."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import dyn_res_pile_manip_b200 as P
from dyn_res_pile_manip_b200 import ops, planner as planner_mod, synthetic
from oracle import pile_oracle as O

N, T, NS = 30, 3, 8


def _install_doubles(pl, W, env, goal):
    goal_t = torch.tensor(goal, dtype=torch.float)

    def evaluate(s0, dens, attr, model_dy, acts, goal_img, coor, w):
        pred = O.rollout(W, 0.08, env.get_cam_extrinsics(), synthetic.GLOBAL_SCALE, s0, dens, attr, acts)
        rew = O.reward_ptcl(pred[:, -1], goal_t, env.get_cam_params(), coor)
        return rew, torch.from_numpy(O.mppi_record(rew.numpy(), acts.numpy(), w)).float()

    pl._mppi_evaluate = evaluate
    pl.device = torch.device("cpu")
    ops.mppi_combine = lambda parts, T_: torch.from_numpy(O.mppi_merge(parts.numpy())).float()


def _plan(world, rank, port, out):
    if world > 1:
        os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    cfg, env = synthetic.default_config(), synthetic.FakeEnv()
    pl = P.PlannerGD(cfg, env)
    W = O.weights_from_seed(0)
    goal = synthetic.make_goal("disc")
    _install_doubles(pl, W, env, goal)
    st, dn = synthetic.make_pile_batch(1, N, seed=2)
    res = pl.trajectory_optimization_mppi(st, dn, np.zeros((1, N), np.float32), goal, None,
                                          synthetic.random_actions(1, T, seed=2)[0], n_sample=NS, n_update_iter=2, seed=5)
    if rank == 0:
        np.save(out, res["action_sequence"])
    if world > 1:
        dist.destroy_process_group()


def _worker(rank, world, port, out):
    _plan(world, rank, port, out)


def test_two_ranks_give_the_single_rank_plan(tmp_path):
    one = str(tmp_path / "one.npy")
    two = str(tmp_path / "two.npy")
    mp.spawn(_worker, args=(1, 0, one), nprocs=1, join=True)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, two), nprocs=2, join=True)
    a, b = np.load(one), np.load(two)
    assert a.shape == (T, 4) and np.isfinite(a).all()
    np.testing.assert_allclose(a, b, rtol=1e-5, atol=1e-6)


def _merge_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(7)
    rew = torch.randn(world, 5, generator=g)
    rew[1, 2] = rew[0, 2]                                   # a tie between ranks: lower trajectory index must win
    idx = torch.tensor([[3, 1, 4, 0, 2], [7, 9, 5, 8, 6]])
    acts = torch.randn(world, 5, 3, 4, generator=g)
    r, i, a = planner_mod.merge_best_across_ranks(rew[rank], idx[rank], acts[rank])
    if rank == 0:
        torch.save((r, i, a, rew, idx, acts), out)
    dist.destroy_process_group()


def test_gd_best_merge_across_two_ranks(tmp_path):
    out = str(tmp_path / "merge.pt")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_merge_worker, args=(2, port, out), nprocs=2, join=True)
    r, i, a, rew, idx, acts = torch.load(out)
    for v in range(5):
        w = 0 if (rew[0, v] > rew[1, v] or (rew[0, v] == rew[1, v] and idx[0, v] < idx[1, v])) else 1
        assert r[v] == rew[w, v] and i[v] == idx[w, v] and torch.equal(a[v], acts[w, v])
    # single process: identity
    r1, i1, a1 = planner_mod.merge_best_across_ranks(rew[0], idx[0], acts[0])
    assert torch.equal(r1, rew[0]) and torch.equal(i1, idx[0])


def test_shard_bounds_partition_the_samples():
    for world in (1, 2, 4, 8):
        spans = [planner_mod.shard_bounds(1024, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == 1024
        assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
    try:
        planner_mod.shard_size(10, 4)
        assert False
    except ValueError:
        pass


def test_record_merge_equals_softmax_on_union():
    rng = np.random.RandomState(0)
    acts, rew = rng.uniform(-4, 4, (64, 5, 1, 4)), rng.uniform(-400, -1, (64, 1))
    recs = [O.mppi_record(rew[i:i + 16, 0], acts[i:i + 16, :, 0], 0.1) for i in range(0, 64, 16)]
    m = O.mppi_merge(recs)
    np.testing.assert_allclose((m[2:] / m[1]).reshape(5, 4), O.mppi_optimize_action(acts, rew, 0.1)[:, 0], rtol=1e-10)
