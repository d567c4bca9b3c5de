#!/usr/bin/env python
"""bench.py -- GNN rollout particle-steps/sec (BASELINE.json metric) on N B200s of one node.

One "step" = one MPPI planner evaluation of the hot path on one batch of synthetic input:
T-step particle-GNN rollout of `samples` sampled action sequences (relation search + propagation
network per horizon step), target-shape reward of the final state, MPPI weighting record, and
-- for N > 1 -- ONE all-gather of the (2+4T)-float record per evaluation.

Two splits are measured in the same run (BASELINE config 3 reads "1024 samples ... sharded over 1/2/4/8"):
  * weak   (headline `value`, "scaling": "weak"): every rank evaluates its own 1024 sequences;
  * strong (`strong` object): 1024 sequences in total, 1024/N per rank.  At N = 1 the same object carries the
    single-GPU times of the 512/256/128-sample shards (what each rank of a 2/4/8-GPU strong split runs; the
    collective is 82 floats) with the per-kernel times that say which kernel stops the scaling.

  python bench.py                       # N=1, BASELINE config 3 @ 300 particles: 1024 x 300 x T=20
  torchrun ... bench.py --gpus 8        # weak + strong split on 8 ranks, NCCL record exchange, + config 5
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
KERNELS = ["nbr_search", "node_encode", "edge_encode", "edge_agg0", "node_update0", "edge_agg1", "node_update1",
           "edge_agg2", "node_update2_predict"]


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


def kernel_rooflines(kms, E, R, B, N, engine_mode, hbm_gbs, bf16_tf):
    """Per-kernel roofline entries from the live per-kernel times (ms) of one model step.
    Algorithmic bytes / flops per launch as stated in DESIGN.md section 5."""
    ce_row = 192 if engine_mode == 2 else H * 4          # tensor engine 2 stores C_e (and P_s) rows as 24-bit words
    ps_row = ce_row
    out = {}

    def hbm(name, nbytes):
        if kms[name] <= 0:
            return
        ach = nbytes / (kms[name] * 1e-3) / 1e9
        out[name] = {"bound": "hbm", "achieved": ach, "peak": hbm_gbs, "unit": "GB/s", "frac": ach / hbm_gbs,
                     "algorithmic_bytes": nbytes, "ms": kms[name]}

    def tensor(name, flops):
        ach = flops / (kms[name] * 1e-3) / 1e12
        out[name] = {"bound": "tensor", "achieved": ach, "peak": bf16_tf, "unit": "TFLOP/s", "frac": ach / bf16_tf,
                     "algorithmic_flops": flops, "ms": kms[name]}

    hbm("nbr_search", R * 24 + 4 * B * (N + 1) + 8 * E + 32 * E)
    hbm("node_encode", R * 16 + R * (3 * H * 4 + ps_row))
    tensor("edge_encode", E * 2 * (6 * H + 3 * H * H))
    for p in range(3):
        hbm("edge_agg%d" % p, E * (ce_row + 4) + R * (H * 4 + ps_row + H * 4) + 4 * (R + B))
    for name in ("node_update0", "node_update1"):
        hbm(name, R * (3 * H * 4) + R * (2 * H * 4 + ps_row))
    hbm("node_update2_predict", R * (3 * H * 4) + R * 24)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg3_n300", choices=sorted(WORKLOADS))
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip plan latency / GD / parity legs (quick kernel runs)")
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
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
            os.environ.pop("NCCL_DEBUG")               # keep NCCL's version banner off stdout: ONE JSON line
        dist.init_process_group("nccl", device_id=dev)

    cfg, env = synthetic.default_config(), synthetic.FakeEnv()
    torch.manual_seed(0)
    model = PropNetDiffDenModel(cfg, True).to(dev)
    planner = PlannerGD(cfg, env)
    goal = synthetic.make_goal("bar")
    lib = _lib.load()
    K, Wm = args.steps, args.warmup
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)     # > 126 MB L2
    st, dn = synthetic.make_pile_batch(1, N, seed=0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed_loop(step_fn, steps, step_fn2=None):
        """Times `steps` calls of step_fn (and, interleaved one for one, of step_fn2: both arms then see the same clock /
        power state -- a B200 settles ~3 % lower after a few hundred ms of sustained load).  -> (ms, t0, t1[, ms2])"""
        evs, evs2 = [], []
        barrier()
        t0 = time.time()
        for i in range(steps):
            for fn, out in ((step_fn, evs), (step_fn2, evs2)):
                if fn is None:
                    continue
                flush.fill_(i & 0xff)                      # evict L2 between timed steps
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                fn(i)
                b.record()
                out.append((a, b))
        barrier()
        t1 = time.time()
        both = [sum(a.elapsed_time(b) for a, b in evs), sum(a.elapsed_time(b) for a, b in evs2)]
        ms = torch.tensor(both, device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        if step_fn2 is None:
            return float(ms[0].item()), t0, t1
        return float(ms[0].item()), t0, t1, float(ms[1].item())

    class Arm:
        """One engine + its action pool + the record exchange of an evaluation."""

        def __init__(self, n_samples, n_particles, horizon, seed0, steps, state):
            self.S, self.N, self.T = n_samples, n_particles, horizon
            self.eng = RolloutEngine(model, planner, n_samples, n_particles, horizon, device=dev, goal=goal,
                                     use_graph=not args.no_graph)
            self.eng.load_state(*state)
            pool = [torch.from_numpy(synthetic.random_actions(n_samples, horizon, seed=seed0 + i)) for i in range(steps)]
            self.pool_dev = [p.to(dev) for p in pool]
            self.pool_host = [p.pin_memory() for p in pool]
            self.rec_all = torch.zeros(world * (2 + 4 * horizon), device=dev)
            self.rec_out = torch.zeros(2 + 4 * horizon, device=dev)
            self.r_host = torch.empty(n_samples, dtype=torch.float32).pin_memory()
            self.rec_host = torch.empty(2 + 4 * horizon, dtype=torch.float32).pin_memory()

        def exchange(self):
            if world > 1:
                dist.all_gather_into_tensor(self.rec_all, self.eng.record)
                _lib.check(lib.pile_mppi_combine(_lib.ptr(self.rec_all), world, self.T, _lib.ptr(self.rec_out),
                                                 ops._stream()), "combine")

        def step_device(self, i):
            self.eng.actions.copy_(self.pool_dev[i % len(self.pool_dev)])
            self.eng.evaluate()
            self.exchange()

        def step_host(self, i):
            self.eng.evaluate_host(self.pool_host[i % len(self.pool_host)], self.r_host, self.rec_host)
            self.exchange()

        def launches_per_step(self):
            return self.eng.launches_per_eval() + (1 if world > 1 else 0)

    # ---- weak split (headline): every rank evaluates `samples` sequences ------------------------------------------
    arm = Arm(samples, N, T, 1000 * rank, K + Wm, (st, dn))
    for i in range(Wm):
        arm.step_device(i)
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    for i in range(Wm):
        arm.step_host(i)
    # device-resident and host-buffer evaluations interleaved one for one inside the same timed region
    total_ms, t0, t2, e2e_ms = timed_loop(lambda i: arm.step_device(Wm + i), K, lambda i: arm.step_host(Wm + i))
    time.sleep(0.2)
    sampler.stop()
    clocks = sampler.summary(t0, t2)
    units = world * samples * N * T
    value = units * K / (total_ms * 1e-3)
    e2e_value = units * K / (e2e_ms * 1e-3)
    launches = 2 * K * arm.launches_per_step()

    def profile_kernels(a, reps=5):
        ms = (_lib.C.c_float * len(KERNELS))()
        s_out = torch.empty(a.S, a.N, 3, device=dev)
        wpack = model.model.packed_weights(dev)
        _lib.check(lib.pile_profile_step(_lib.ptr(wpack), _lib.ptr(a.eng.attr), _lib.ptr(a.eng.dens), _lib.ptr(a.eng.s0),
                                         _lib.ptr(a.pool_dev[0]), a.T * 4, planner.pusher.ref(), 0.08, a.S, a.N,
                                         _lib.ptr(a.eng.scratch), _lib.ptr(s_out), reps, ms, ops._stream()),
                   "pile_profile_step")
        return dict(zip(KERNELS, [float(v) for v in ms]))

    # ---- strong split: 1024 sequences in total -------------------------------------------------------------------
    strong = None
    if args.workload.startswith("cfg3"):
        weak_ms_per_step = total_ms / K
        if world > 1 and samples % world == 0:
            per = samples // world
            sarm = Arm(per, N, T, 5000 + 1000 * rank, K + Wm, (st, dn))
            for i in range(Wm):
                sarm.step_device(i)
            s_ms, _, _ = timed_loop(lambda i: sarm.step_device(Wm + i), K)
            launches += K * sarm.launches_per_step()
            strong = {"total_samples": samples, "samples_per_gpu": per, "n_gpus": world,
                      "value": samples * N * T * K / (s_ms * 1e-3), "unit": UNIT, "ms_per_step": s_ms / K,
                      "n1_ms_per_step": weak_ms_per_step,
                      "efficiency_vs_n1": (weak_ms_per_step / world) / (s_ms / K),
                      "note": "n1_ms_per_step = this run's %d-sample evaluation per GPU (the weak leg, max over ranks, "
                              "incl. its record exchange)" % samples}
            if rank == 0:
                strong["kernel_ms"] = profile_kernels(sarm)
            del sarm
        elif world == 1:
            shards = []
            for g in (2, 4, 8):
                per = samples // g
                sarm = Arm(per, N, T, 5000 + g, K + Wm, (st, dn))
                for i in range(Wm):
                    sarm.step_device(i)
                s_ms, _, _ = timed_loop(lambda i: sarm.step_device(Wm + i), K)
                launches += K * sarm.launches_per_step()
                shards.append({"n_gpus": g, "samples_per_gpu": per, "ms_per_step": s_ms / K,
                               "projected_value": samples * N * T / (s_ms / K * 1e-3),
                               "projected_efficiency": (weak_ms_per_step / g) / (s_ms / K),
                               "kernel_ms": profile_kernels(sarm)})
                del sarm
            strong = {"total_samples": samples, "n_gpus": 1, "value": value, "unit": UNIT, "ms_per_step": weak_ms_per_step,
                      "efficiency_vs_n1": 1.0, "single_gpu_shards": shards,
                      "note": "single_gpu_shards: the per-rank work of a 2/4/8-GPU strong split timed on this one GPU "
                              "(the collective of a multi-GPU run is 82 floats); projected_efficiency = (t_1024 / g) / t_shard"}

    # ---- config 5 on the 8-rank run: 16384 x 300 x T=30 over 8 GPUs -------------------------------------------------
    cfg5 = None
    if world == 8 and args.workload == "cfg3_n300":
        s5, n5, t5 = WORKLOADS["cfg5"]
        k5 = max(3, min(K, 5))
        a5 = Arm(s5, n5, t5, 9000 + 1000 * rank, k5 + Wm, (st, dn))
        for i in range(Wm):
            a5.step_device(i)
        ms5, _, _ = timed_loop(lambda i: a5.step_device(Wm + i), k5)
        launches += k5 * a5.launches_per_step()
        cfg5 = {"workload": "cfg5: %d samples x %d particles x T=%d over 8 GPUs + NCCL MPPI record exchange" % (8 * s5, n5, t5),
                "value": 8 * s5 * n5 * t5 * k5 / (ms5 * 1e-3), "unit": UNIT, "ms_per_step": ms5 / k5, "steps": k5}
        del a5

    # ---- rank 0: roofline, parity curve, plan latency, GD refinement, CPU baseline ------------------------------
    line = None
    if rank == 0:
        hbm_gbs, bf16_tf, _, peak_kind = peaks()
        mode = lib.pile_get_tensor_cores()
        kms = profile_kernels(arm)
        rel = ops.relations_from_buffer(arm.eng.scratch, False, samples, N)
        E = int(rel.n_rel.sum().item())
        R = samples * N
        per_kernel = kernel_rooflines(kms, E, R, samples, N, mode, hbm_gbs, bf16_tf)
        step_ms = sum(kms.values())
        # dominant kernel = largest share of the model step summed over its launches (k_edge_agg and the particle
        # update launch three times per step); `achieved` is per launch (average over the launches)
        groups = {"nbr_search": ["nbr_search"], "node_encode": ["node_encode"], "edge_encode": ["edge_encode"],
                  "edge_agg": ["edge_agg0", "edge_agg1", "edge_agg2"],
                  "node_update": ["node_update0", "node_update1", "node_update2_predict"]}
        share = {g: sum(kms[k] for k in ks) / step_ms for g, ks in groups.items()}
        dom = max(share, key=share.get)
        members = [k for k in groups[dom] if k in per_kernel]
        first = per_kernel[members[0]]
        avg_ms = sum(kms[k] for k in members) / len(members)
        work = first.get("algorithmic_bytes", first.get("algorithmic_flops"))
        ach = work / (avg_ms * 1e-3) / (1e9 if first["bound"] == "hbm" else 1e12)
        roof = {"kernel": dom, "bound": first["bound"], "achieved": ach, "peak": first["peak"], "unit": first["unit"],
                "frac": ach / first["peak"], "traffic": None, "launches_per_model_step": len(members),
                "avg_launch_ms": avg_ms,
                "peak_source": peak_kind + (" copy" if first["bound"] == "hbm" else " bf16 cuBLAS burst")}
        tr = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.isfile(tr):
            tj = json.load(open(tr))
            roof["traffic"] = tj.get(dom)
        roof["kernel_ms"] = kms
        roof["kernel_share_of_model_step"] = share
        roof["kernels"] = per_kernel
        roof["E_relations"] = E
        f_ref, f_alg = O.flops_per_sample_step(N, E / samples, H)
        roof["alg_tflops_whole_step"] = f_alg * samples / (step_ms * 1e-3) / 1e12
        roof["note"] = ("tensor kernels: tcgen05 bf16 hi/lo split, 3 tensor passes per algorithmic MAC (tensor-pipe MACs = 3x achieved)"
                        if mode else "FP32 CUDA-core tile GEMM engine (parity anchor)")

        extras = world == 1 and not args.no_extras
        # BASELINE config 3 is a resolution SWEEP (the regressor's output range): the other particle counts, device-timed
        # the same way (L2 flushed, CUDA events, K steps after W warm-ups)
        sweep = None
        if extras and args.workload == "cfg3_n300":
            sweep = {}
            for wl in ("cfg3_n50", "cfg3_n100", "cfg3_n200"):
                s_w, n_w, t_w = WORKLOADS[wl]
                st_w, dn_w = synthetic.make_pile_batch(1, n_w, seed=0)
                a_w = Arm(s_w, n_w, t_w, 7000 + n_w, K + Wm, (st_w, dn_w))
                for i in range(Wm):
                    a_w.step_device(i)
                ms_w, _, _ = timed_loop(lambda i: a_w.step_device(Wm + i), K)
                launches += K * a_w.launches_per_step()
                sweep[wl] = {"value": s_w * n_w * t_w * K / (ms_w * 1e-3), "unit": UNIT, "ms_per_step": ms_w / K,
                             "kernel_ms": profile_kernels(a_w)}
                del a_w
        parity = parity_report(model, planner, dev) if extras else None
        plan = plan_latency(model, planner, env, goal, dev, args) if extras else None
        cpu = None
        torch_cuda = None
        if world == 1 and not args.no_cpu_baseline:
            # the reference algorithm with torch's own CUDA kernels on this B200 (dense one-hot bmm path), chunk sized
            # so the dense [chunk, 10N, N] one-hot tensors use a few GB
            gchunk = 256 if N >= 200 else 512
            try:
                r_gpu, t_gpu, _ = cpu_reference_rate(N, T, gchunk, 3.0, warmup=1, device="cuda")
                torch_cuda = {"value": r_gpu, "unit": UNIT, "kind": "oracle port on torch-CUDA (cuBLAS/ATen), same B200",
                              "sample": "%d passes of %d action sequences x %d particles x T=%d" % (len(t_gpu), gchunk, N, T)}
            except RuntimeError as err:      # e.g. out of memory for the dense [B, 10N, N] tensors
                torch_cuda = {"value": None, "error": str(err)[:120]}
            torch.cuda.empty_cache()
            chunk = 16 if N >= 200 else 32
            rate, times, cores = cpu_reference_rate(N, T, chunk, 12.0)
            cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": "%d passes of %d of %d action sequences x %d particles x T=%d (oracle/pile_oracle.py, "
                             "torch CPU, dense one-hot form)" % (len(times), chunk, samples, N, T)}

        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
                "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": args.workload, "samples_per_gpu": samples, "particles": N, "horizon": T,
                           "nf_effect": H, "relations_per_particle": E / R, "cuda_graph": not args.no_graph,
                           "gemm_engine": {0: "fp32 CUDA cores", 1: "tcgen05 (smem A)", 2: "tcgen05 (TMEM A)"}[mode],
                           "cache": "L2 flushed (256 MB fill) between timed steps; per-step intermediates (%.0f MB) exceed L2"
                                    % (E * H * 4 / 1e6),
                           "collective": "all_gather of %d floats per evaluation" % (2 + 4 * T) if world > 1 else "none"},
                "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms / K,
                        "h2d_bytes_per_step": samples * T * 4 * 4,
                        "d2h_bytes_per_step": samples * 4 + (2 + 4 * T) * 4},
                "gpu_launches": launches,
                "clocks": clocks, "roofline": roof}
        if sweep:
            line["resolution_sweep"] = sweep
        if strong:
            line["strong"] = strong
        if cfg5:
            line["cfg5"] = cfg5
        if parity:
            line["parity"] = parity
        if plan:
            line["mpc_plan_latency"] = plan
        if cpu:
            line["cpu_baseline"] = cpu
        if torch_cuda:
            line["torch_cuda_reference"] = torch_cuda
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if line:
        print(json.dumps(line))


def parity_report(model, planner, dev):
    """Drift over the full horizon against vectors the real reference produced at BASELINE's sizes
    (tests/golden/golden_big_v1.npz; SURVEY 8d 'Parity report'): per-step ||d||/||s|| and relation-set Jaccard."""
    from dyn_res_pile_manip_b200 import ops
    path = os.path.join(ROOT, "tests", "golden", "golden_big_v1.npz")
    base = os.path.join(ROOT, "tests", "golden", "golden_v1.npz")
    if not (os.path.isfile(path) and os.path.isfile(base)):
        return None
    g, b = np.load(path), np.load(base)
    weights = {k[2:]: torch.from_numpy(v) for k, v in b.items() if k.startswith("w/")}
    keep = {k: v.detach().clone() for k, v in model.state_dict().items()}

    def pair_set(c):
        c = np.asarray(c).astype(np.int64)
        return set((c[:, 0] << 40 | c[:, 1] << 20 | c[:, 2]).tolist())

    def coo(rel):
        return np.concatenate([np.concatenate([np.full((e.shape[0], 1), i), e], axis=1) for i, e in enumerate(rel.edge_sets())])

    out = {"source": "tests/golden/golden_big_v1.npz (real reference, CPU, seed-0 weights)"}
    try:
        for tag, name in (("K", "n300_T20_tamed"), ("H", "n300_T20_random_init"), ("J", "n100_T10_tamed")):
            nb, ns, N, T = [int(v) for v in g[tag + "/dims"]]
            w = {k: v.clone() for k, v in weights.items()}
            for k in ("model.particle_predictor.linear_1.weight", "model.particle_predictor.linear_1.bias"):
                w[k] = w[k] * float(g[tag + "/tame"])
            model.load_state_dict(w)
            planner.particle_num = N
            acts = torch.from_numpy(g[tag + "/acts"]).to(dev)
            with torch.no_grad():
                pred = planner.ptcl_model_rollout(torch.from_numpy(g[tag + "/s0"]).to(dev), torch.from_numpy(g[tag + "/dens"]).to(dev),
                                                  torch.zeros(nb, N, device=dev), model, acts)["model_rollout"]["state_pred"]
                ref = torch.from_numpy(g[tag + "/state_pred"]).to(dev)
                drift = [float((pred[:, t] - ref[:, t]).double().norm() / ref[:, t].double().norm()) for t in range(T)]
                jac = []
                s = torch.from_numpy(g[tag + "/s0"]).to(dev).repeat(ns, 1, 1)
                for t in range(T):
                    sd = planner.gen_s_delta(s, acts[:, t].contiguous())
                    mine, want = pair_set(coo(ops.build_relations(s, sd, 0.08))), pair_set(g[tag + "/rel%d" % t])
                    jac.append(len(mine & want) / len(mine | want))
                    s = pred[:, t].contiguous()
            out[name] = {"one_step_rel": drift[0], "drift_T": drift, "jaccard_T": jac, "samples": nb * ns, "particles": N,
                         "horizon": T, "predictor_scale": float(g[tag + "/tame"])}
    finally:
        model.load_state_dict(keep)
    return out


def plan_latency(model, planner, env, goal, dev, args):
    """MPC plan latency (BASELINE.json metric, second half) and the GD refinement (config 4)."""
    from dyn_res_pile_manip_b200 import PropNetDiffDenModel, synthetic
    from dyn_res_pile_manip_b200.engine import RolloutEngine
    # one MPPI planner evaluation of BASELINE config 2 (256 samples x 100 particles x T=10) through the host-buffer
    # call, >= 50 timed calls after 5 warm-ups
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
    del eng2
    # the same evaluation through the planner's MPPI entry (numpy state / goal / mean sequence in -> numpy plan out:
    # noise sampling on the host, captured engine underneath), 3 MPPI iterations per call
    mean2 = synthetic.random_actions(1, t2, seed=3)[0]
    ml = []
    for i in range(25):
        t_a = time.perf_counter()
        planner.trajectory_optimization_mppi(st2, dn2, np.zeros((1, n2), np.float32), goal, model, mean2, n_sample=s2,
                                             n_update_iter=3, seed=i)
        ml.append((time.perf_counter() - t_a) * 1e3)
    ml = sorted(ml[5:])
    plan["mppi_planner"] = {"p50_ms": ml[len(ml) // 2], "calls": len(ml), "iterations": 3,
                            "workload": "trajectory_optimization_mppi: 256 samples x 100 particles x T=10, 3 iterations, "
                                        "numpy in -> numpy plan out"}

    # the reference's own MPC entry point with its shipped configuration (config/mpc/config.yaml:38-43,
    # env/flex_env.py:1020): 50 trajectories x 30 state variants, 100 particles, horizon 1, time budget
    # 2000 ms -> 27 Adam iterations (planners.py:679-682)
    st3, dn3 = synthetic.make_pile_batch(30, 100, seed=0)
    act3 = synthetic.random_actions(50, 1, seed=9).transpose(1, 0, 2).astype(np.float64)
    gd = []
    for i in range(12):
        t_a = time.perf_counter()
        res = planner.trajectory_optimization_ptcl_multi_traj(
            st3, dn3, np.zeros((30, 100), np.float32), goal, model, act3, np.zeros(1), 50, 1, 200, None, None,
            time_lim=2000)
        gd.append((time.perf_counter() - t_a) * 1e3)
    gd = sorted(gd[2:])
    plan["gd_planner"] = {"p50_ms": gd[len(gd) // 2], "calls": len(gd), "iterations": int(res["iter_num"]) + 1,
                          "optim_loop_ms": res["times"]["rollout_time"] + res["times"]["optim_time"],
                          "workload": "trajectory_optimization_ptcl_multi_traj: 50 traj x 30 variants x 100 particles, "
                                      "T=1, numpy in -> result dict out (reference budget for this call: 2000 ms)"}

    # BASELINE config 4: gradient-based refinement, forward + backward through T=20, 128 samples x 300 particles,
    # one Adam iteration = rollout with tape + last-step reward | reward/rollout backward + Adam + clamp
    st4, dn4 = synthetic.make_pile_batch(1, 300, seed=0)
    act4 = synthetic.random_actions(128, 20, seed=4).transpose(1, 0, 2).astype(np.float64)
    iters = 6
    best = None
    for rep in range(3):
        res4 = planner.trajectory_optimization_ptcl_multi_traj(
            st4, dn4, np.zeros((1, 300), np.float32), goal, model, act4, np.zeros(20), 128, 20, iters, None, None,
            rollout_best_action_sequence=False)
        cur = (res4["times"]["rollout_time"] / iters, res4["times"]["optim_time"] / iters)
        if best is None or sum(cur) < sum(best):
            best = cur
    plan["cfg4_gd_refinement"] = {"fwd_ms": best[0], "bwd_ms": best[1], "iter_ms": best[0] + best[1],
                                  "value": 128 * 300 * 20 / ((best[0] + best[1]) * 1e-3), "unit": UNIT,
                                  "workload": "cfg4: 128 samples x 300 particles x T=20, one Adam iteration (captured graphs): "
                                              "fwd = rollout with tape + reward + best tracking, bwd = reward/rollout backward + Adam/clamp"}
    planner._gd_loops.clear()

    # many concurrent MPC calls (the Bayesian-optimisation data generation runs five episodes per candidate resolution,
    # data_gen/res_rgr_data.py:128-221): five scenes of the shipped MPC call planned one after the other vs in one
    # batched loop (trajectory_optimization_ptcl_multi_scene)
    scenes = [synthetic.make_pile_batch(30, 100, seed=60 + k) for k in range(5)]
    attr5 = np.zeros((30, 100), np.float32)
    seq_ms, bat_ms = [], []
    for rep in range(5):
        t_a = time.perf_counter()
        for st_k, dn_k in scenes:
            planner.trajectory_optimization_ptcl_multi_traj(st_k, dn_k, attr5, goal, model, act3, np.zeros(1), 50, 1, 200,
                                                            None, None, time_lim=2000)
        seq_ms.append((time.perf_counter() - t_a) * 1e3)
        t_a = time.perf_counter()
        planner.trajectory_optimization_ptcl_multi_scene([a_ for a_, _ in scenes], [b_ for _, b_ in scenes], [attr5] * 5, goal,
                                                         model, act3, np.zeros(1), 50, 1, 200, time_lim=2000)
        bat_ms.append((time.perf_counter() - t_a) * 1e3)
    plan["bo_concurrent_planning"] = {"scenes": 5, "one_by_one_ms": sorted(seq_ms[1:])[len(seq_ms[1:]) // 2],
                                      "batched_ms": sorted(bat_ms[1:])[len(bat_ms[1:]) // 2],
                                      "workload": "5 x (50 traj x 30 variants x 100 particles, T=1, 27 Adam iterations): five "
                                                  "trajectory_optimization_ptcl_multi_traj calls vs one ..._multi_scene call"}
    planner._gd_loops.clear()

    # resolution regressor: infer_param = host cv2 planes + H2D + captured 15-launch graph + D2H;
    # device part alone timed with CUDA events around graph replays (457 MB weight stream)
    from dyn_res_pile_manip_b200 import MPCResRgrNoPool
    torch.manual_seed(5)
    rgr = MPCResRgrNoPool({"train_res_cls": {"state_h": 224, "state_w": 224, "res_dim": 6}})
    st6, _ = synthetic.make_pile_batch(1, 200, seed=4)
    fg6 = (synthetic.render_observation(st6[0], env)[..., -1] / env.global_scale < 0.599 / 0.8).astype(np.float32)
    gm6 = (goal < 0.5).astype(np.float32)
    rl = []
    for i in range(14):
        t_a = time.perf_counter()
        rgr.infer_param(fg6, gm6)
        rl.append((time.perf_counter() - t_a) * 1e3)
    g6 = rgr._graph
    evs = []
    for i in range(10):
        flush_rgr = torch.empty(192 * 1024 * 1024, dtype=torch.uint8, device=dev).fill_(i)      # evict L2
        a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a_.record(); g6['graph'].replay(); b_.record()
        evs.append((a_, b_))
    torch.cuda.synchronize()
    dev_ms = sorted(a_.elapsed_time(b_) for a_, b_ in evs)[len(evs) // 2]
    plan["resolution_regressor"] = {"infer_param_p50_ms": sorted(rl[4:])[len(rl[4:]) // 2], "device_graph_ms": dev_ms,
                                    "weight_stream_gbs": 114193217 * 4 / (dev_ms * 1e-3) / 1e9,
                                    "workload": "MPCResRgrNoPool.infer_param: two 720x720 masks -> particle count (cv2 planes on "
                                                "the host, 5 conv + 5 linear layers on the device, batch 1, 457 MB of weights)"}
    del rgr, g6

    # training step: the reference's loop body (train/train_gnn_dyn.py:150-199) on a padded variable-N
    # batch: 3 roll-out steps forward, MSE, backward with all 18 weight gradients, Adam
    import torch.nn.functional as F
    torch.manual_seed(0)
    tmodel = PropNetDiffDenModel(synthetic.default_config(), True).to(dev)
    opt = torch.optim.Adam(tmodel.parameters(), lr=1e-4)
    Bt, Nt, n_roll = 32, 300, 3
    st7, dn7 = synthetic.make_pile_batch(Bt, Nt, seed=7)
    rng7 = np.random.RandomState(7)
    nums7 = rng7.randint(150, Nt + 1, size=Bt)
    states7 = torch.tensor(np.stack([st7 + rng7.normal(0, 0.003, st7.shape).astype(np.float32) * k for k in range(n_roll + 1)], 1)).to(dev)
    sdel7 = torch.tensor((rng7.normal(0, 0.01, (Bt, n_roll, Nt, 3)) * (rng7.uniform(size=(Bt, n_roll, Nt, 1)) < 0.3)).astype(np.float32)).to(dev)
    attr7, dens7, nums_t = torch.zeros(Bt, Nt, device=dev), torch.tensor(dn7).to(dev), torch.tensor(nums7)
    tl = []
    for i in range(8):
        torch.cuda.synchronize()
        t_a = time.perf_counter()
        opt.zero_grad()
        s_cur, loss = states7[:, 0], 0.
        for t in range(n_roll):
            s_pred = tmodel.predict_one_step(attr7, s_cur, sdel7[:, t], dens7, nums_t)
            for j in range(Bt):
                loss = loss + F.mse_loss(s_pred[j, :nums7[j]], states7[j, t + 1, :nums7[j]])
            s_cur = s_pred
        (loss / (n_roll * Bt)).backward()
        opt.step()
        torch.cuda.synchronize()
        tl.append((time.perf_counter() - t_a) * 1e3)
    plan["training_step"] = {"p50_ms": sorted(tl[2:])[len(tl[2:]) // 2], "particle_steps_per_s": Bt * Nt * n_roll / (sorted(tl[2:])[len(tl[2:]) // 2] * 1e-3),
                             "workload": "32 samples x <=300 particles (padded, particle_nums), 3 roll-out steps: forward + "
                                         "per-sample MSE + backward (18 weight gradients) + torch Adam, through predict_one_step"}
    del tmodel, opt

    # a checkpoint with another hidden width (model/gnn_dyn.py:119; north_star: "hidden ~150") runs on the general-width
    # engine (csrc/general.cu, FP32 CUDA-core block GEMMs): one model step of config 2's batch, forward only and
    # forward + input/weight gradients
    import copy
    cfg_w = copy.deepcopy(synthetic.default_config())
    cfg_w['train']['particle']['nf_effect'] = 150
    torch.manual_seed(0)
    wmodel = PropNetDiffDenModel(cfg_w, True).to(dev)
    s8 = torch.tensor(st2).to(dev).repeat(s2, 1, 1)
    sd8 = (0.01 * torch.randn(s8.shape, device=dev, generator=torch.Generator(device=dev).manual_seed(8)))
    a8, d8 = torch.zeros(s2, n2, device=dev), torch.tensor(dn2).to(dev).repeat(s2)
    fw, fb = [], []
    for i in range(8):
        torch.cuda.synchronize()
        t_a = time.perf_counter()
        with torch.no_grad():
            wmodel.predict_one_step(a8, s8, sd8, d8)
        torch.cuda.synchronize()
        fw.append((time.perf_counter() - t_a) * 1e3)
        t_a = time.perf_counter()
        wmodel.zero_grad()
        wmodel.predict_one_step(a8, s8, sd8, d8).square().sum().backward()
        torch.cuda.synchronize()
        fb.append((time.perf_counter() - t_a) * 1e3)
    plan["general_width"] = {"nf_effect": 150, "forward_ms": sorted(fw[2:])[len(fw[2:]) // 2],
                             "forward_backward_ms": sorted(fb[2:])[len(fb[2:]) // 2],
                             "particle_steps_per_s": s2 * n2 / (sorted(fw[2:])[len(fw[2:]) // 2] * 1e-3),
                             "workload": "one model step, 256 samples x 100 particles, nf_effect = 150 (padded to 192) on the "
                                         "general-width engine: forward (no_grad: hoisted inference form, Hp-wide layers on tcgen05); forward + backward with all 18 "
                                         "weight gradients (FP32 CUDA cores)"}
    del wmodel

    # the other planner-side piece of an MPC step: RGB-D observation -> 30 particle
    # re-samplings (env/flex_env.py:933-951), host observation in -> host particles out
    from dyn_res_pile_manip_b200 import observation as OBS
    st5, _ = synthetic.make_pile_batch(1, 300, seed=0)
    obs5 = synthetic.render_observation(st5[0], env)
    ol = []
    for i in range(12):
        t_a = time.perf_counter()
        OBS.obs2ptcl_fixed_num_batch(obs5, 100, 30, env.get_cam_params(), env.global_scale, seed=i)
        ol.append((time.perf_counter() - t_a) * 1e3)
    ol = sorted(ol[2:])
    plan["obs_to_particles"] = {"p50_ms": ol[len(ol) // 2], "calls": len(ol),
                                "workload": "obs2ptcl_fixed_num_batch: 720x720 RGB-D -> 30 x 100 particles "
                                            "(depth2fgpcd, 1 cm voxel downsample, FPS, recenter), numpy in -> numpy out"}
    if not args.no_cpu_baseline:
        from oracle import obs_oracle as OO
        t_a = time.perf_counter()
        depth5 = obs5[..., -1] / env.global_scale
        for i in range(2):      # the reference repeats all four stages for each of the 30 re-samplings
            fg = OO.voxel_down_sample(OO.depth2fgpcd(depth5, depth5 < 0.599 / 0.8, env.get_cam_params()), 0.01)
            pk, r_ = OO.fps(fg, 100, i)
            OO.recenter(fg, pk, r=min(0.02, 0.5 * r_))
        plan["obs_to_particles"]["cpu_port_ms"] = (time.perf_counter() - t_a) * 1e3 / 2 * 30
        plan["obs_to_particles"]["cpu_sample"] = "2 of 30 re-samplings timed (oracle/obs_oracle.py, numpy), scaled to 30"
    return plan


if __name__ == "__main__":
    main()
