"""Where does the host-buffer evaluation spend its extra time?  Per-iteration CUDA-event and wall times of
RolloutEngine.evaluate (device actions) and evaluate_host (pinned host actions in, host reward + record out)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dyn_res_pile_manip_b200 import PlannerGD, PropNetDiffDenModel, synthetic
from dyn_res_pile_manip_b200.engine import RolloutEngine
cfg, env = synthetic.default_config(), synthetic.FakeEnv()
torch.manual_seed(0)
model = PropNetDiffDenModel(cfg, True).cuda()
planner = PlannerGD(cfg, env)
S, N, T = 1024, 300, 20
eng = RolloutEngine(model, planner, S, N, T, goal=synthetic.make_goal("bar"))
st, dn = synthetic.make_pile_batch(1, N, seed=0)
eng.load_state(st, dn)
acts = [torch.from_numpy(synthetic.random_actions(S, T, seed=i)) for i in range(4)]
dev = [a.cuda() for a in acts]; pin = [a.pin_memory() for a in acts]
r_host = torch.empty(S).pin_memory(); rec_host = torch.empty(2 + 4 * T).pin_memory()
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
for i in range(3):
    eng.actions.copy_(dev[i]); eng.evaluate()
torch.cuda.synchronize()
def run(kind, n=8, do_flush=True):
    out = []
    for i in range(n):
        if do_flush:
            flush.fill_(i)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        a.record()
        if kind == "dev":
            eng.actions.copy_(dev[i % 4]); eng.evaluate()
        else:
            eng.evaluate_host(pin[i % 4], r_host, rec_host)
        b.record()
        torch.cuda.synchronize()
        out.append((a.elapsed_time(b), (time.perf_counter() - t0) * 1e3))
    return out
for kind in ("dev", "host", "dev", "host"):
    r = run(kind)
    print(kind, "event ms:", " ".join("%.2f" % e for e, _ in r), "| wall ms:", " ".join("%.2f" % w for _, w in r))
