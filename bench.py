#!/usr/bin/env python
"""bench.py -- GNN rollout particle-steps/sec (BASELINE.json metric) on N B200s of one node.

One "step" = one MPPI planner evaluation of the hot path on one batch of synthetic input:
T-step particle-GNN rollout of `samples` sampled action sequences (relation search + propagation
network per horizon step), target-shape reward of the final state, MPPI weighting record, and
-- for N > 1 -- ONE all-gather of the (2+4T)-float record per evaluation.  Work per GPU is fixed
(weak scaling: every rank evaluates its own `samples` sequences), value = all ranks' particle-steps
/ max-over-ranks device time.

  python bench.py                       # N=1, BASELINE config 3 @ 300 particles: 1024 x 300 x T=20
  torchrun ... bench.py --gpus 8        # same per-GPU batch on 8 ranks + NCCL record exchange
  python bench.py --impl reference      # the reference algorithm's CPU path (oracle port) on host cores
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "gnn_rollout_particle_steps_per_sec"
UNIT = "particle-steps/s"
H = 64
WORKLOADS = {   # name -> (samples per GPU, particles, horizon)   (BASELINE.json configs)
    "cfg3_n300": (1024, 300, 20),
    "cfg3_n200": (1024, 200, 20),
    "cfg3_n100": (1024, 100, 20),
    "cfg3_n50": (1024, 50, 20),
    "cfg2": (256, 100, 10),
    "cfg5": (2048, 300, 30),
    "cfg1": (1, 100, 1),          # the reference's own CPU-runnable case (reference arm / parity only)
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return d["hbm_gbs"], d["bf16_tflops"], d.get("sm_max_mhz"), "measured"
    return 6650.0, 1590.0, 1965.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self):
        if self.proc:
            self.proc.terminate()

    def summary(self, t0, t1):
        rows = [r for t, r in self.rows if t0 <= t <= t1 and len(r) >= 9] or [r for _, r in self.rows if len(r) >= 9]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        reasons = set()
        for r in rows:
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(float(r[1]) for r in rows), "sm_max_mhz": float(rows[0][2]),
                "power_w_max": max(float(r[3]) for r in rows), "samples": len(rows), "reasons": sorted(reasons)}


def cpu_reference_rate(N, T, chunk, min_seconds, warmup=0, steps=None, device="cpu"):
    """Time the oracle (dense one-hot formulation = the reference's own algorithm) on a bounded sample of the
    workload: `chunk` action sequences x N particles x T steps per pass (rollout + last-step reward).
    device="cpu": the reference's CPU path on all host cores; device="cuda": the same torch code on the B200
    (cuBLAS/ATen kernels), reported as the 'existing GPU path' comparator."""
    from dyn_res_pile_manip_b200 import synthetic
    from oracle import pile_oracle as O
    torch.set_num_threads(os.cpu_count())
    env = synthetic.FakeEnv()
    W = {k: v.to(device) for k, v in O.weights_from_seed(0).items()}
    st, dn = synthetic.make_pile_batch(1, N, seed=0)
    goal = synthetic.make_goal("bar")
    coords = np.argwhere(goal < 0.5)[:, ::-1].astype(np.float32)
    coor, _ = synthetic.fps_np(coords, min(5 * N, len(coords)), 0)
    goal_t, coor_t = torch.from_numpy(goal), torch.from_numpy(coor).to(device)
    s0, dens, attr = torch.from_numpy(st).to(device), torch.from_numpy(dn).to(device), torch.zeros(1, N, device=device)

    def one_pass(seed):
        acts = torch.from_numpy(synthetic.random_actions(chunk, T, seed=seed)).to(device)
        with torch.no_grad():
            pred = O.rollout(W, 0.08, env.get_cam_extrinsics(), synthetic.GLOBAL_SCALE, s0, dens, attr, acts)
            O.reward_ptcl(pred[:, -1], goal_t, env.get_cam_params(), coor_t)
        if device != "cpu":
            torch.cuda.synchronize()

    for w in range(warmup):
        one_pass(100 + w)
    times = []
    t_begin = time.perf_counter()
    k = 0
    while True:
        t0 = time.perf_counter()
        one_pass(k)
        times.append(time.perf_counter() - t0)
        k += 1
        if steps is not None:
            if k >= steps:
                break
        elif time.perf_counter() - t_begin >= min_seconds:
            break
    per_pass = chunk * N * T
    return per_pass * len(times) / sum(times), times, torch.get_num_threads()


def run_reference(args, samples, N, T):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    chunk = 1 if args.workload == "cfg1" else (16 if N >= 200 else 32)
    rate, times, cores = cpu_reference_rate(N, T, chunk, 0, warmup=min(args.warmup, 1),
                                            steps=max(1, min(args.steps, 50 if args.workload == "cfg1" else 3)))
    sample = "%d of %d action sequences x %d particles x T=%d per step (oracle/pile_oracle.py, dense one-hot form)" % (
        chunk, samples, N, T)
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": len(times),
            "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "samples_per_gpu": samples, "particles": N, "horizon": T,
                       "nf_effect": H, "bounded_sample": sample},
            "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg3_n300", choices=sorted(WORKLOADS))
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    samples, N, T = WORKLOADS[args.workload]

    if args.impl == "reference":
        run_reference(args, samples, N, T)
        return

    import torch.distributed as dist
    from dyn_res_pile_manip_b200 import PlannerGD, PropNetDiffDenModel, _lib, ops, synthetic
    from dyn_res_pile_manip_b200.engine import RolloutEngine
    from oracle import pile_oracle as O   # flop model + CPU baseline leg only

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch N>1 with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    cfg, env = synthetic.default_config(), synthetic.FakeEnv()
    torch.manual_seed(0)
    model = PropNetDiffDenModel(cfg, True).to(dev)
    planner = PlannerGD(cfg, env)
    goal = synthetic.make_goal("bar")
    eng = RolloutEngine(model, planner, samples, N, T, device=dev, goal=goal, use_graph=not args.no_graph)
    st, dn = synthetic.make_pile_batch(1, N, seed=0)
    eng.load_state(st, dn)

    K, Wm = args.steps, args.warmup
    pool = [torch.from_numpy(synthetic.random_actions(samples, T, seed=1000 * rank + i)) for i in range(K + Wm)]
    pool_dev = [p.to(dev) for p in pool]
    pool_host = [p.pin_memory() for p in pool]
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)     # > 126 MB L2
    rec_all = torch.zeros(world * (2 + 4 * T), device=dev)
    rec_out = torch.zeros(2 + 4 * T, device=dev)

    def exchange():
        if world > 1:
            dist.all_gather_into_tensor(rec_all, eng.record)
            lib = _lib.load()
            _lib.check(lib.pile_mppi_combine(_lib.ptr(rec_all), world, T, _lib.ptr(rec_out), ops._stream()), "combine")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed_loop(step_fn):
        evs = []
        barrier()
        t0 = time.time()
        for i in range(K):
            flush.fill_(i & 0xff)                      # evict L2 between timed steps
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            step_fn(Wm + i)
            b.record()
            evs.append((a, b))
        barrier()
        t1 = time.time()
        ms = torch.tensor([sum(a.elapsed_time(b) for a, b in evs)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), t0, t1

    # ---- device-resident arm ------------------------------------------------------------------------------
    def step_device(i):
        eng.actions.copy_(pool_dev[i])
        eng.evaluate()
        exchange()

    for i in range(Wm):
        step_device(i)
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    total_ms, t0, t1 = timed_loop(step_device)

    # ---- end-to-end arm: host buffers in, host results out, through the engine's public call ----------
    r_host = torch.empty(samples, dtype=torch.float32).pin_memory()
    rec_host = torch.empty(2 + 4 * T, dtype=torch.float32).pin_memory()

    def step_host(i):
        eng.evaluate_host(pool_host[i], r_host, rec_host)
        exchange()

    for i in range(Wm):
        step_host(i)
    e2e_ms, _, t2 = timed_loop(step_host)
    time.sleep(0.2)
    sampler.stop()
    clocks = sampler.summary(t0, t2)

    units = world * samples * N * T
    value = units * K / (total_ms * 1e-3)
    e2e_value = units * K / (e2e_ms * 1e-3)

    # ---- roofline of the dominant kernel, timed live with CUDA events (rank 0) --------------------------
    line = None
    if rank == 0:
        hbm_gbs, bf16_tf, _, peak_kind = peaks()
        lib = _lib.load()
        ms6 = (_lib.C.c_float * 6)()
        s_out = torch.empty(samples, N, 3, device=dev)
        wpack = model.model.packed_weights(dev)
        _lib.check(lib.pile_profile_step(_lib.ptr(wpack), _lib.ptr(eng.attr), _lib.ptr(eng.dens), _lib.ptr(eng.s0),
                                         _lib.ptr(pool_dev[0]), T * 4, planner.pusher.ref(), 0.08, samples, N,
                                         _lib.ptr(eng.scratch),
                                         _lib.ptr(s_out), 5, ms6, ops._stream()), "pile_profile_step")
        names = ["nbr_search", "node_encode", "edge_encode", "propagate0", "propagate1", "propagate2_predict"]
        kms = dict(zip(names, [float(v) for v in ms6]))
        rel = ops.relations_from_buffer(eng.scratch, False, samples, N)
        E = int(rel.n_rel.sum().item())
        R = samples * N
        flops = {"edge_encode": E * 2 * (6 * H + 3 * H * H),
                 "node_encode": R * 2 * (5 * H + 4 * H * H),
                 "propagate0": R * 2 * (3 * H * H) + E * 3 * H, "propagate1": R * 2 * (3 * H * H) + E * 3 * H,
                 "propagate2_predict": R * 2 * (2 * H * H + 3 * H) + E * 3 * H}
        ce_row = 192 if lib.pile_get_tensor_cores() == 2 else H * 4     # tensor engine 2 stores C_e as 24-bit words
        hbm_bytes = {"nbr_search": R * 24 + 4 * (samples * (N + 1)) + 8 * E + 32 * E,
                     "edge_encode": E * (32 + ce_row),
                     "propagate0": E * (ce_row + 4) + R * H * 4 * 7}
        dom = max(kms, key=kms.get)
        step_ms = sum(kms.values())
        if dom in ("edge_encode", "node_encode"):
            ach = flops[dom] / (kms[dom] * 1e-3) / 1e12
            roof = {"kernel": dom, "bound": "tensor", "achieved": ach, "peak": bf16_tf, "unit": "TFLOP/s",
                    "frac": ach / bf16_tf, "traffic": None, "peak_source": peak_kind + " bf16 cuBLAS burst",
                    "note": ("tcgen05 bf16 hi/lo split, 3 tensor passes per algorithmic MAC (tensor-pipe MACs = 3x achieved)"
                             if lib.pile_get_tensor_cores() else "FP32 CUDA-core tile GEMM engine (parity anchor)")}
        else:
            by = hbm_bytes.get(dom, hbm_bytes["propagate0"])
            ach = by / (kms[dom] * 1e-3) / 1e9
            roof = {"kernel": dom, "bound": "hbm", "achieved": ach, "peak": hbm_gbs, "unit": "GB/s",
                    "frac": ach / hbm_gbs, "traffic": None, "peak_source": peak_kind + " copy"}
        tr = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.isfile(tr):
            roof["traffic"] = json.load(open(tr)).get(dom)
        roof["kernel_ms"] = kms
        roof["kernel_share_of_model_step"] = {k: v / step_ms for k, v in kms.items()}
        roof["E_relations"] = E
        f_ref, f_alg = O.flops_per_sample_step(N, E / samples, H)
        roof["alg_tflops_whole_step"] = f_alg * samples / (step_ms * 1e-3) / 1e12

        # MPC plan latency (BASELINE.json metric, second half): one MPPI planner evaluation of BASELINE config 2
        # (256 samples x 100 particles x T=10) through the host-buffer call, >= 50 timed calls after 5 warm-ups
        plan = None
        if world == 1:
            s2, n2, t2 = WORKLOADS["cfg2"]
            eng2 = RolloutEngine(model, planner, s2, n2, t2, device=dev, goal=goal, use_graph=not args.no_graph)
            st2, dn2 = synthetic.make_pile_batch(1, n2, seed=0)
            eng2.load_state(st2, dn2)
            acts2 = [torch.from_numpy(synthetic.random_actions(s2, t2, seed=50 + i)).pin_memory() for i in range(8)]
            r2 = torch.empty(s2, dtype=torch.float32).pin_memory()
            rec2 = torch.empty(2 + 4 * t2, dtype=torch.float32).pin_memory()
            lat = []
            for i in range(55):
                t_a = time.perf_counter()
                eng2.evaluate_host(acts2[i % 8], r2, rec2)
                lat.append((time.perf_counter() - t_a) * 1e3)
            lat = sorted(lat[5:])
            plan = {"p50_ms": lat[len(lat) // 2], "p90_ms": lat[int(len(lat) * 0.9)], "calls": len(lat),
                    "workload": "cfg2: 256 samples x 100 particles x T=10, host actions in -> host reward + MPPI record out"}

            # the reference's own MPC entry point with its shipped configuration (config/mpc/config.yaml:38-43,
            # env/flex_env.py:1020): 50 trajectories x 30 state variants, 100 particles, horizon 1, time budget
            # 2000 ms -> 27 Adam iterations (planners.py:679-682)
            st3, dn3 = synthetic.make_pile_batch(30, 100, seed=0)
            act3 = synthetic.random_actions(50, 1, seed=9).transpose(1, 0, 2).astype(np.float64)
            gd = []
            for i in range(8):
                t_a = time.perf_counter()
                res = planner.trajectory_optimization_ptcl_multi_traj(
                    st3, dn3, np.zeros((30, 100), np.float32), goal, model, act3, np.zeros(1), 50, 1, 200, None, None,
                    time_lim=2000)
                gd.append((time.perf_counter() - t_a) * 1e3)
            gd = sorted(gd[2:])
            plan["gd_planner"] = {"p50_ms": gd[len(gd) // 2], "calls": len(gd), "iterations": int(res["iter_num"]) + 1,
                                  "workload": "trajectory_optimization_ptcl_multi_traj: 50 traj x 30 variants x 100 particles, "
                                              "T=1, numpy in -> result dict out (reference budget for this call: 2000 ms)"}

            # the other planner-side piece of an MPC step: RGB-D observation -> 30 particle
            # re-samplings (env/flex_env.py:933-951), host observation in -> host particles out
            from dyn_res_pile_manip_b200 import observation as OBS
            st4, _ = synthetic.make_pile_batch(1, 300, seed=0)
            obs4 = synthetic.render_observation(st4[0], env)
            ol = []
            for i in range(12):
                t_a = time.perf_counter()
                OBS.obs2ptcl_fixed_num_batch(obs4, 100, 30, env.get_cam_params(), env.global_scale, seed=i)
                ol.append((time.perf_counter() - t_a) * 1e3)
            ol = sorted(ol[2:])
            plan["obs_to_particles"] = {"p50_ms": ol[len(ol) // 2], "calls": len(ol),
                                        "workload": "obs2ptcl_fixed_num_batch: 720x720 RGB-D -> 30 x 100 particles "
                                                    "(depth2fgpcd, 1 cm voxel downsample, FPS, recenter), numpy in -> numpy out"}
            if not args.no_cpu_baseline:
                from oracle import obs_oracle as OO
                t_a = time.perf_counter()
                depth4 = obs4[..., -1] / env.global_scale
                for i in range(2):      # the reference repeats all four stages for each of the 30 re-samplings
                    fg4 = OO.voxel_down_sample(OO.depth2fgpcd(depth4, depth4 < 0.599 / 0.8, env.get_cam_params()), 0.01)
                    pk4, r4 = OO.fps(fg4, 100, i)
                    OO.recenter(fg4, pk4, r=min(0.02, 0.5 * r4))
                plan["obs_to_particles"]["cpu_port_ms"] = (time.perf_counter() - t_a) * 1e3 / 2 * 30
                plan["obs_to_particles"]["cpu_sample"] = "2 of 30 re-samplings timed (oracle/obs_oracle.py, numpy), scaled to 30"

        cpu = None
        torch_cuda = None
        if world == 1 and not args.no_cpu_baseline:
            chunk = 16 if N >= 200 else 32
            # the reference algorithm with torch's own CUDA kernels on this B200 (dense one-hot bmm path)
            try:
                r_gpu, t_gpu, _ = cpu_reference_rate(N, T, chunk, 3.0, warmup=1, device="cuda")
                torch_cuda = {"value": r_gpu, "unit": UNIT, "kind": "oracle port on torch-CUDA (cuBLAS/ATen), same B200",
                              "sample": "%d passes of %d action sequences x %d particles x T=%d" % (len(t_gpu), chunk, N, T)}
            except RuntimeError as err:      # e.g. out of memory for the dense [B, 10N, N] tensors
                torch_cuda = {"value": None, "error": str(err)[:120]}
            rate, times, cores = cpu_reference_rate(N, T, chunk, 12.0)
            cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": "%d passes of %d of %d action sequences x %d particles x T=%d (oracle/pile_oracle.py, "
                             "torch CPU, dense one-hot form)" % (len(times), chunk, samples, N, T)}

        launches = eng.launches_per_eval() + (1 if world > 1 else 0) + 1   # + D2D action copy is a memcpy, + flush fill
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
                "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": args.workload, "samples_per_gpu": samples, "particles": N, "horizon": T,
                           "nf_effect": H, "relations_per_particle": E / R, "cuda_graph": not args.no_graph,
                           "cache": "L2 flushed (256 MB fill) between timed steps; per-step intermediates (%.0f MB) exceed L2"
                                    % (E * H * 4 / 1e6),
                           "collective": "all_gather of %d floats per evaluation" % (2 + 4 * T) if world > 1 else "none"},
                "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms / K,
                        "h2d_bytes_per_step": samples * T * 4 * 4,
                        "d2h_bytes_per_step": samples * 4 + (2 + 4 * T) * 4},
                "gpu_launches": (eng.launches_per_eval() + (1 if world > 1 else 0)) * K,
                "clocks": clocks, "roofline": roof}
        if plan:
            line["mpc_plan_latency"] = plan
        if cpu:
            line["cpu_baseline"] = cpu
        if torch_cuda:
            line["torch_cuda_reference"] = torch_cuda
        del launches
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if line:
        print(json.dumps(line))


if __name__ == "__main__":
    main()
