"""Cycle trace of one warp of k_edge_encode_tc (measurement aid; run on the GPU box).
The trace points are compiled in only with -DPILE_ENABLE_TRACE: build a variant library, e.g.
  cp -r dyn_res_pile_manip_b200/csrc abl/trace && make -C abl/trace FLAGS+=-DPILE_ENABLE_TRACE
and run with PILE_GNN_LIB=abl/trace/libpilegnn.so."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dyn_res_pile_manip_b200 import PlannerGD, PropNetDiffDenModel, _lib, ops, synthetic
from dyn_res_pile_manip_b200.engine import RolloutEngine

cfg, env = synthetic.default_config(), synthetic.FakeEnv()
torch.manual_seed(0)
model = PropNetDiffDenModel(cfg, True).cuda()
planner = PlannerGD(cfg, env)
which = int(sys.argv[1]) if len(sys.argv) > 1 else 0
ops.set_tensor_cores(2 if which == 2 else 1)
eng = RolloutEngine(model, planner, 1024, 300, 1, use_graph=False)
st, dn = synthetic.make_pile_batch(1, 300, seed=0)
eng.load_state(st, dn)
eng.actions.copy_(torch.from_numpy(synthetic.random_actions(1024, 1, seed=1)))
eng.evaluate(); torch.cuda.synchronize()
buf = torch.zeros(4096, dtype=torch.int64, device="cuda")
lib = _lib.load()
which = int(sys.argv[1]) if len(sys.argv) > 1 else 0
_lib.check(lib.pile_debug_set_trace(_lib.ptr(buf), 4096, which), "trace")
eng.evaluate(); torch.cuda.synchronize()
_lib.check(lib.pile_debug_set_trace(None, 0, which), "trace")
v = buf.cpu().numpy()
v = v[v != 0]
tags = (v >> 56) & 0xff
clk = v & ((1 << 56) - 1)
names = {1: "tile start", 2: "layer top", 3: "fences done", 4: "barrier passed", 5: "mma done", 6: "tile end", 7: "kernel entry", 8: "prologue done", 9: "weights landed", 10: "loop done"}
t0 = clk[0]
prev = t0
for i, (tg, c) in enumerate(zip(tags, clk)):
    if i > 140:
        break
    print("%4d %-16s +%6d  (t=%d)" % (i, names.get(int(tg), tg), c - prev, c - t0))
    prev = c
