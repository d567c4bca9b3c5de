import sys, time, copy
sys.path.insert(0, '/root/repo')
import torch, numpy as np
from dyn_res_pile_manip_b200 import PropNetDiffDenModel, synthetic
for nf in (150, 96, 256):
    cfg = copy.deepcopy(synthetic.default_config()); cfg['train']['particle']['nf_effect'] = nf
    torch.manual_seed(0)
    m = PropNetDiffDenModel(cfg, True).cuda()
    st, dn = synthetic.make_pile_batch(1, 100, seed=0)
    s = torch.tensor(st).cuda().repeat(256, 1, 1); sd = 0.01 * torch.randn_like(s)
    a, d = torch.zeros(256, 100, device='cuda'), torch.tensor(dn).cuda().repeat(256)
    for mode in ('no_grad', 'grad'):
        ts = []
        for i in range(8):
            torch.cuda.synchronize(); t = time.perf_counter()
            if mode == 'no_grad':
                with torch.no_grad(): m.predict_one_step(a, s, sd, d)
            else:
                m.requires_grad_(False); m.predict_one_step(a, s.clone().requires_grad_(True), sd, d)
            torch.cuda.synchronize(); ts.append((time.perf_counter() - t) * 1e3)
        print(nf, mode, "%.2f ms" % sorted(ts[2:])[3])
