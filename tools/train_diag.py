import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import dyn_res_pile_manip_b200 as P
from dyn_res_pile_manip_b200 import synthetic
from oracle import pile_oracle as O
from test_gpu_training import training_loss
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
g = np.load(os.path.join(ROOT, "tests/golden/golden_train_v1.npz"))
b = np.load(os.path.join(ROOT, "tests/golden/golden_v1.npz"))
W = {k[2:]: torch.from_numpy(v) for k, v in b.items() if k.startswith("w/")}
model = P.PropNetDiffDenModel(synthetic.default_config(), True); model.load_state_dict(W); model = model.cuda()
args = [torch.tensor(g[k]).cuda() for k in ("states", "states_delta", "attrs", "dens")]
for n_roll in (1, 2):
    model.zero_grad()
    a2 = [args[0][:, :n_roll + 1], args[1][:, :n_roll], args[2], args[3]]
    loss = training_loss(model, *a2, torch.tensor(g["particle_nums"]))
    loss.backward()
    Wc = {k: v.clone().requires_grad_(True) for k, v in W.items()}
    lo = O.training_loss(Wc, 0.08, torch.tensor(g["states"][:, :n_roll + 1]), torch.tensor(g["states_delta"][:, :n_roll]), torch.tensor(g["attrs"]), torch.tensor(g["dens"]), torch.tensor(g["particle_nums"]))
    lo.backward()
    print("n_roll", n_roll, "loss", loss.item(), lo.item())
    for k, p in model.named_parameters():
        ref = Wc[k].grad.numpy()
        got = p.grad.cpu().numpy()
        err = np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-12)
        extra = ""
        if got.ndim == 2 and got.shape[1] <= 8:
            extra = " percol " + " ".join("%.1e" % (np.abs(got[:, c] - ref[:, c]).max() / max(np.abs(ref).max(), 1e-12)) for c in range(got.shape[1]))
        print("  %-45s %.2e  max|ref| %.3e%s" % (k, err, np.abs(ref).max(), extra))
