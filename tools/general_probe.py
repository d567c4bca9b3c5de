import sys, copy
sys.path.insert(0, '/root/repo')
import torch
from dyn_res_pile_manip_b200 import PropNetDiffDenModel, synthetic
nf = int(sys.argv[1]) if len(sys.argv) > 1 else 150
cfg = copy.deepcopy(synthetic.default_config()); cfg['train']['particle']['nf_effect'] = nf
torch.manual_seed(0)
m = PropNetDiffDenModel(cfg, True).cuda()
st, dn = synthetic.make_pile_batch(1, 100, seed=0)
s = torch.tensor(st).cuda().repeat(256, 1, 1); sd = 0.01 * torch.randn_like(s)
a, d = torch.zeros(256, 100, device='cuda'), torch.tensor(dn).cuda().repeat(256)
with torch.no_grad():
    for _ in range(2): m.predict_one_step(a, s, sd, d)
torch.cuda.synchronize()
