"""Evaluation time (captured graph, L2 flushed) and a digest of states + MPPI record for A/B runs of two libraries:
    [PILE_GNN_LIB=...] python tools/eval_time.py [particles] [horizon] [samples ...]"""
import hashlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dyn_res_pile_manip_b200 import PlannerGD, PropNetDiffDenModel, synthetic
from dyn_res_pile_manip_b200.engine import RolloutEngine

N = int(sys.argv[1]) if len(sys.argv) > 1 else 300
T = int(sys.argv[2]) if len(sys.argv) > 2 else 20
sizes = [int(v) for v in sys.argv[3:]] or [128, 1024]
cfg, env = synthetic.default_config(), synthetic.FakeEnv()
torch.manual_seed(0)
model = PropNetDiffDenModel(cfg, True).cuda()
planner = PlannerGD(cfg, env)
goal = synthetic.make_goal("bar")
st, dn = synthetic.make_pile_batch(1, N, seed=0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
out = []
for S in sizes:
    eng = RolloutEngine(model, planner, S, N, T, goal=goal)
    eng.load_state(st, dn)
    eng.actions.copy_(torch.from_numpy(synthetic.random_actions(S, T, seed=1)).cuda())
    for _ in range(3):
        eng.evaluate()
    torch.cuda.synchronize()
    ms = []
    for i in range(10):
        flush.fill_(i)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); eng.evaluate(); b.record(); torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    digest = hashlib.sha1(eng.states.cpu().numpy().tobytes() + eng.record.cpu().numpy().tobytes()).hexdigest()[:10]
    out.append("%d x %d x T=%d: %.3f ms (%s)" % (S, N, T, sorted(ms)[len(ms) // 2], digest))
    del eng
print("[%s]" % os.path.basename(os.path.dirname(os.environ.get("PILE_GNN_LIB", "./default"))), " | ".join(out))
