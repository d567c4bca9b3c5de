"""Experiment: one 1024-row evaluation vs. two 512-row evaluations on two streams (measurement aid)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dyn_res_pile_manip_b200 import PlannerGD, PropNetDiffDenModel, _lib, ops, synthetic
from dyn_res_pile_manip_b200.engine import RolloutEngine

B, N, T = 1024, 300, 20
cfg, env = synthetic.default_config(), synthetic.FakeEnv()
torch.manual_seed(0)
model = PropNetDiffDenModel(cfg, True).cuda()
planner = PlannerGD(cfg, env)
st, dn = synthetic.make_pile_batch(1, N, seed=0)
acts = torch.from_numpy(synthetic.random_actions(B, T, seed=1)).cuda()
lib = _lib.load()
if len(sys.argv) > 1 and hasattr(lib, "pile_set_sm_limit"):
    lib.pile_set_sm_limit(int(sys.argv[1]))

def timeit(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

full = RolloutEngine(model, planner, B, N, T, use_graph=False)
full.load_state(st, dn); full.actions.copy_(acts)
print("1 x %d rows: %.2f ms" % (B, timeit(full.evaluate)))
for parts in (2, 3, 4):
    rows = B // parts // 4 * 4
    engs = [RolloutEngine(model, planner, rows, N, T, use_graph=False) for _ in range(parts)]
    streams = [torch.cuda.Stream() for _ in range(parts)]
    for k, e in enumerate(engs):
        e.load_state(st, dn); e.actions.copy_(acts[k * rows:(k + 1) * rows])
    def run():
        cur = torch.cuda.current_stream()
        for e, s in zip(engs, streams):
            s.wait_stream(cur)
            with torch.cuda.stream(s):
                e.evaluate()
        for s in streams:
            cur.wait_stream(s)
    ms = timeit(run)
    print("%d x %d rows on %d streams: %.2f ms (scaled to %d rows: %.2f ms)" % (parts, rows, parts, ms, B, ms * B / (rows * parts)))
