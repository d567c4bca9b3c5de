"""Per-kernel-bucket times of one model step (pile_profile_step) for the library named by PILE_GNN_LIB, default the
in-tree build.  Used for the ablation measurements quoted in DESIGN.md section 5: a variant library is built with
one group of loads / stores / GEMMs compiled out (temporary `#if` guards around them, never committed) and this
script prints the six bucket times for it."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dyn_res_pile_manip_b200 import PlannerGD, PropNetDiffDenModel, _lib, ops, synthetic
from dyn_res_pile_manip_b200.engine import RolloutEngine

B, N = int(os.environ.get("ABL_B", 1024)), int(os.environ.get("ABL_N", 300))
cfg, env = synthetic.default_config(), synthetic.FakeEnv()
torch.manual_seed(0)
model = PropNetDiffDenModel(cfg, True).cuda()
planner = PlannerGD(cfg, env)
eng = RolloutEngine(model, planner, B, N, 1, use_graph=False)
st, dn = synthetic.make_pile_batch(1, N, seed=0)
eng.load_state(st, dn)
eng.actions.copy_(torch.from_numpy(synthetic.random_actions(B, 1, seed=1)))
eng.evaluate(); torch.cuda.synchronize()
lib = _lib.load()
ms6 = (_lib.C.c_float * 9)()
s_out = torch.empty(B, N, 3, device="cuda")
wpack = model.model.packed_weights(torch.device("cuda"))
_lib.check(lib.pile_profile_step(_lib.ptr(wpack), _lib.ptr(eng.attr), _lib.ptr(eng.dens), _lib.ptr(eng.s0),
                                 _lib.ptr(eng.actions), 4, planner.pusher.ref(), 0.08, B, N, _lib.ptr(eng.scratch),
                                 _lib.ptr(s_out), 10, ms6, ops._stream()), "pile_profile_step")
print(os.path.basename(os.environ.get("PILE_GNN_LIB", "default")), " ".join("%.1f" % (1e3 * v) for v in ms6), "us  sum %.1f" % (1e3 * sum(ms6)))
