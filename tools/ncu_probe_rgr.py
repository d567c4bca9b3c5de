"""Regressor forward (batch 1) + one training step for ncu launch lists."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch.nn.functional as F
from dyn_res_pile_manip_b200 import MPCResRgrNoPool, PropNetDiffDenModel, synthetic
what = sys.argv[1] if len(sys.argv) > 1 else "rgr"
if what == "rgr":
    torch.manual_seed(5)
    rgr = MPCResRgrNoPool({"train_res_cls": {"state_h": 224, "state_w": 224, "res_dim": 6}})
    x = torch.rand(1, 6, 224, 224).cuda()
    for _ in range(2):
        y = rgr.forward(x)
    torch.cuda.synchronize()
else:
    torch.manual_seed(0)
    m = PropNetDiffDenModel(synthetic.default_config(), True).cuda()
    Bt, Nt = 32, 300
    st, dn = synthetic.make_pile_batch(Bt, Nt, seed=7)
    s = torch.tensor(st).cuda(); sd = torch.tensor(np.random.RandomState(0).normal(0, 0.01, st.shape).astype(np.float32)).cuda()
    for _ in range(2):
        out = m.predict_one_step(torch.zeros(Bt, Nt).cuda(), s, sd, torch.tensor(dn).cuda())
        (out ** 2).mean().backward()
    torch.cuda.synchronize()
