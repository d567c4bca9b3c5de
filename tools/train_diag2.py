import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import torch.nn.functional as F
import dyn_res_pile_manip_b200 as P
from dyn_res_pile_manip_b200 import synthetic, ops, _lib
from oracle import pile_oracle as O
g = np.load(os.path.join(ROOT, "tests/golden/golden_train_v1.npz"))
b = np.load(os.path.join(ROOT, "tests/golden/golden_v1.npz"))
W = {k[2:]: torch.from_numpy(v) for k, v in b.items() if k.startswith("w/")}
model = P.PropNetDiffDenModel(synthetic.default_config(), True); model.load_state_dict(W); model = model.cuda()
dev = torch.device("cuda")
s = torch.tensor(g["states"][:, 0]).cuda(); sd = torch.tensor(g["states_delta"][:, 0]).cuda()
a = torch.tensor(g["attrs"][:, 0]).cuda(); dn = torch.tensor(g["dens"]).cuda()
nums = torch.tensor(g["particle_nums"]).int().cuda()
B, N, _ = s.shape
lib = _lib.load()
wpack = model.model.packed_weights(dev)
tape = torch.zeros(lib.pile_train_tape_bytes(B, N), dtype=torch.uint8, device=dev)
out = ops.train_forward_raw(wpack, a, dn, s, sd, 0.08, nums, tape)
gp = torch.randn(B, N, 3, device=dev, generator=torch.Generator(device=dev).manual_seed(1))
grads = torch.zeros(lib.pile_train_grad_offset(18), device=dev)
scratch = torch.zeros(lib.pile_train_scratch_bytes(B, N), dtype=torch.uint8, device=dev)
g_s, g_sd = ops.train_backward_raw(wpack, dn, tape, B, N, gp, grads, scratch)
torch.cuda.synchronize()
R, E, H = B * N, B * 10 * N, 64
def up(x): return (x + 255) // 256 * 256
off = 0
def take(nfl):
    global off
    v = scratch[off:off + nfl * 4].view(torch.float32); off += up(nfl * 4); return v
gA, gEff, gZ, gP, gAgg, gH0 = [take(R * H).view(R, H) for _ in range(6)]
dX0 = take(R * 8).view(R, 8)
gR3, dZr, dZs, gR2 = [take(E * H).view(E, H) for _ in range(4)]
rel = ops.relations_from_buffer(tape, 2, B, N)
# ---- reference with retained grads (dense, CPU)
Wc = {k: v.clone().requires_grad_(True) for k, v in W.items()}
sc, sdc, ac, dc = s.cpu(), sd.cpu(), a.cpu(), dn.cpu()
adj = O.adjacency(sc, sdc, 0.08, g["particle_nums"])
Rr, Rs = O.one_hot_relations(adj)
n_rel = Rr.shape[1]
d = dc / 5000.; d_node = d.view(B, 1, 1).expand(B, N, 1); d_rel = d.view(B, 1, 1).expand(B, n_rel, 1)
lin = lambda p, x: F.linear(x, Wc[p + ".weight"], Wc[p + ".bias"])
x = torch.cat([sdc, ac.unsqueeze(-1), d_node], 2)
p_enc = torch.relu(lin("model.particle_encoder.model.2", torch.relu(lin("model.particle_encoder.model.0", x))))
y = torch.cat([Rr.bmm(ac.unsqueeze(-1)), Rs.bmm(ac.unsqueeze(-1)), Rr.bmm(sc) - Rs.bmm(sc), d_rel], 2)
r1 = torch.relu(lin("model.relation_encoder.model.0", y)); r2 = torch.relu(lin("model.relation_encoder.model.2", r1))
r_enc = torch.relu(lin("model.relation_encoder.model.4", r2)); r_enc.retain_grad(); r2.retain_grad()
eff = p_enc
for _ in range(3):
    z = torch.cat([r_enc, Rr.bmm(eff), Rs.bmm(eff), d_rel], 2)
    e_rel = torch.relu(lin("model.relation_propagator.linear", z))
    agg = Rr.transpose(1, 2).bmm(e_rel)
    eff = torch.relu(lin("model.particle_propagator.linear", torch.cat([p_enc, agg, d_node], 2)) + eff)
pred = lin("model.particle_predictor.linear_1", torch.relu(lin("model.particle_predictor.linear_0", eff))) + sc
print("forward max diff", float((pred.detach() - out.cpu()).abs().max()))
(pred * gp.cpu()).sum().backward()
ne = rel.n_rel.cpu().numpy()
for bb in range(B):
    mine = gR3[bb * 10 * N: bb * 10 * N + ne[bb]].cpu()
    ref = r_enc.grad[bb, :ne[bb]]
    mine2 = gR2[bb * 10 * N: bb * 10 * N + ne[bb]].cpu()
    ref2 = r2.grad[bb, :ne[bb]]
    err = (mine - ref).abs().max(1).values
    err2 = (mine2 - ref2).abs().max(1).values
    print("sample", bb, "edges", ne[bb], "gR3 max err %.2e (scale %.2e) rows>1e-6: %s" % (float(err.max()), float(ref.abs().max()), torch.nonzero(err > 1e-6 * float(ref.abs().max()) * 100).flatten()[:12].tolist()),
          "| gR2 max err %.2e (scale %.2e) rows: %s" % (float(err2.max()), float(ref2.abs().max()), torch.nonzero(err2 > 1e-4 * float(ref2.abs().max())).flatten()[:12].tolist()))
# tape layout: csr (6 int arrays) then X0, H0, P, eff x3, agg x3, Q, Y0, R1, R2, R3, M x3
offt = 0
def ttake(nbytes):
    global offt
    v = tape[offt:offt + nbytes]; offt += up(nbytes); return v
for n in (B * (N + 1), E, E, B * (N + 1), E, E):
    ttake(n * 4)
X0 = ttake(R * 8 * 4); nodes = [ttake(R * H * 4) for _ in range(9)]
Y0 = ttake((E + 128) * 8 * 4).view(torch.float32).view(-1, 8)
R1t, R2t, R3t = [ttake(E * H * 4).view(torch.float32).view(E, H) for _ in range(3)]
e = 148
mine, ref = R3t[e].cpu(), r_enc[0, e].detach()
diffmask = (mine > 0) != (ref > 0)
print("row", e, "mask differs at", torch.nonzero(diffmask).flatten().tolist(), "mine", mine[diffmask].tolist(), "ref", ref[diffmask].tolist())
print("Y0 row", Y0[e].cpu().tolist(), "ref y", y[0, e].tolist())
z = F.linear(r2[0, e], Wc["model.relation_encoder.model.4.weight"], Wc["model.relation_encoder.model.4.bias"]).detach()
print("pre-activation at those k (cpu):", z[diffmask].tolist())
