"""All 18 weight gradients of the general-width engine vs the oracle's autograd for several widths (diagnostic)."""
import copy, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dyn_res_pile_manip_b200 as P
from dyn_res_pile_manip_b200 import synthetic
from oracle import pile_oracle as O

for nf, B, N in [(64, 2, 30), (96, 2, 30), (128, 2, 30), (150, 2, 30), (150, 3, 100), (150, 2, 31), (150, 4, 37), (192, 2, 30), (256, 2, 30), (1, 2, 20), (63, 2, 20), (65, 2, 20)]:
    cfg = copy.deepcopy(synthetic.default_config()); cfg['train']['particle']['nf_effect'] = nf
    torch.manual_seed(2)
    model = P.PropNetDiffDenModel(cfg, True).cuda()
    model.model.planner_engines = False
    st, dn = synthetic.make_pile_batch(B, N, seed=3)
    sd = np.random.RandomState(3).normal(0, 0.01, st.shape).astype(np.float32)
    W = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in model.state_dict().items()}
    ref = O.predict_one_step(W, 0.08, torch.zeros(B, N), torch.tensor(st), torch.tensor(sd), torch.tensor(dn))
    out = model.predict_one_step(torch.zeros(B, N).cuda(), torch.tensor(st).cuda(), torch.tensor(sd).cuda(), torch.tensor(dn).cuda())
    w = torch.randn(ref.shape, generator=torch.Generator().manual_seed(1))
    (out * w.cuda()).sum().backward(); (ref * w).sum().backward()
    errs = []
    for k, p in model.named_parameters():
        g = W[k].grad.numpy()
        errs.append(np.abs(p.grad.cpu().numpy() - g).max() / max(np.abs(g).max(), 1e-12))
    print("nf %3d B %d N %3d fwd %.1e | " % (nf, B, N, (out.detach().cpu() - ref.detach()).abs().max()), " ".join("%.0e" % e for e in errs))
