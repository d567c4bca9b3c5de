"""Per-kernel times of one model step + a checksum of a T=2 rollout, for A/B runs under different environment switches
(e.g. PILE_AGG_RING=0 / 1):  python tools/ab_step.py [samples particles]"""
import hashlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dyn_res_pile_manip_b200 import PlannerGD, PropNetDiffDenModel, _lib, ops, synthetic
from dyn_res_pile_manip_b200.engine import RolloutEngine

B, N = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1024, 300)
cfg, env = synthetic.default_config(), synthetic.FakeEnv()
torch.manual_seed(0)
model = PropNetDiffDenModel(cfg, True).cuda()
planner = PlannerGD(cfg, env)
eng = RolloutEngine(model, planner, B, N, 2, use_graph=False)
st, dn = synthetic.make_pile_batch(1, N, seed=0)
if os.environ.get("AB_SORT"):          # particles in a spatially coherent index order (8 x 8 cells, row-major)
    import numpy as np
    p = st[0]
    q = np.floor((p[:, :2] - p[:, :2].min(0)) / (np.ptp(p[:, :2], axis=0) + 1e-9) * 8).clip(0, 7).astype(int)
    st = st[:, np.lexsort((q[:, 0], q[:, 1]))]
eng.load_state(st, dn)
eng.actions.copy_(torch.from_numpy(synthetic.random_actions(B, 2, seed=1)))
lib = _lib.load()
eng.evaluate(); torch.cuda.synchronize()
digest = hashlib.sha1(eng.states.cpu().numpy().tobytes()).hexdigest()[:16]
names = ["nbr", "node_enc", "edge_enc", "agg0", "upd0", "agg1", "upd1", "agg2", "upd2"]
ms = (_lib.C.c_float * 9)()
s_out = torch.empty(B, N, 3, device="cuda")
wpack = model.model.packed_weights(torch.device("cuda"))
_lib.check(lib.pile_profile_step(_lib.ptr(wpack), _lib.ptr(eng.attr), _lib.ptr(eng.dens), _lib.ptr(eng.s0), _lib.ptr(eng.actions), 8,
                                 planner.pusher.ref(), 0.08, B, N, _lib.ptr(eng.scratch), _lib.ptr(s_out), 10, ms, ops._stream()), "profile")
env_s = " ".join("%s=%s" % (k, v) for k, v in os.environ.items() if k.startswith("PILE_"))
print("[%s] %dx%d" % (env_s, B, N), " ".join("%s %.1f" % (n, 1e3 * m) for n, m in zip(names, ms)), "| sum %.1f us | states sha1 %s" % (1e3 * sum(ms), digest))
