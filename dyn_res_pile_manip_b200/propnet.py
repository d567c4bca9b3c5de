"""Drop-in replacement of the reference dynamics model (model/gnn_dyn.py).

Same constructor, same `state_dict` keys (so reference checkpoints load with
`load_state_dict(torch.load(...), strict=False)`, visualize_mpc.py:37-40), same
`predict_one_step(a_cur, s_cur, s_delta, particle_dens, particle_nums=None)` and
`model.forward(a_cur, s_cur, s_delta, Rr, Rs, particle_dens)` signatures -- but every
tensor op runs in libpilegnn's sm_100a kernels.  `Rr`/`Rs` may be the reference's dense
one-hot matrices or (preferred) one `ops.Relations` object passed as `Rr` with `Rs=None`.

Differentiable w.r.t. `s_cur` and `s_delta` (the only gradients the planner needs, planners.py:674) through the
dgrad-only kernels; when gradients are enabled and a parameter requires grad (the training loop,
train/train_gnn_dyn.py:150-199) the step runs through the training kernels (csrc/train.cu) and returns the
gradients of all 18 tensors as well.  `model.requires_grad_(False)` selects the dgrad-only path.
"""
import torch
import torch.nn as nn

from . import _lib, ops


class _Stack(nn.Module):
    """Linear layers registered under the names the reference checkpoint uses."""

    def __init__(self, seq_name, dims):
        super().__init__()
        if seq_name is None:
            return
        layers = []
        for i, (fan_in, fan_out) in enumerate(dims):
            layers.append(nn.Linear(fan_in, fan_out))
            layers.append(nn.ReLU())
        setattr(self, seq_name, nn.Sequential(*layers))


class _Named(nn.Module):
    def __init__(self, **linears):
        super().__init__()
        for name, (fan_in, fan_out) in linears.items():
            setattr(self, name, nn.Linear(fan_in, fan_out))


class _StepFn(torch.autograd.Function):
    """One model step; relation lists either searched (rel=None) or supplied."""

    @staticmethod
    def forward(ctx, s_cur, s_delta, attr, dens, owner, rel, particle_nums):
        need_grad = s_cur.requires_grad or s_delta.requires_grad
        s_cur_c, s_delta_c = ops._f32(s_cur.detach()), ops._f32(s_delta.detach())
        ops._require_cuda(s_cur_c, "s_cur")
        dev = s_cur_c.device
        B, N, _ = s_cur_c.shape
        attr_c, dens_c = ops._f32(attr.detach(), dev), ops._f32(dens.detach(), dev)
        wpack = owner.packed_weights(dev)
        scratch = owner.workspace.scratch(B, N, dev)
        tape = ops.new_tape(B, N, 1, dev) if need_grad else None
        if rel is None:
            pn = None
            if particle_nums is not None:
                pn = torch.as_tensor(particle_nums).to(device=dev, dtype=torch.int32).contiguous()
            out = ops.predict_step_raw(wpack, attr_c, dens_c, s_cur_c, s_delta_c, owner.adj_thresh, pn, scratch, tape)
            owner.last_relations_buffer = (tape if need_grad else scratch, need_grad, B, N)
        else:
            out = ops.forward_relations_raw(wpack, attr_c, dens_c, s_cur_c, s_delta_c, rel, scratch, tape)
        ctx.owner, ctx.tape, ctx.dims, ctx.wpack, ctx.dens = owner, tape, (B, N), wpack, dens_c
        return out

    @staticmethod
    def backward(ctx, g):
        B, N = ctx.dims
        g = ops._f32(g)
        g_s, g_sd = ops.step_backward_raw(ctx.wpack, ctx.dens, ctx.tape, B, N, g,
                                          ctx.owner.workspace.bwd(B, N, g.device))
        return g_s, g_sd, None, None, None, None, None


class _TrainStepFn(torch.autograd.Function):
    """One model step with weight gradients (relations always searched; `particle_nums` masks padded particles)."""

    @staticmethod
    def forward(ctx, s_cur, s_delta, attr, dens, owner, particle_nums, rel, *params):
        s_cur_c, s_delta_c = ops._f32(s_cur.detach()), ops._f32(s_delta.detach())
        ops._require_cuda(s_cur_c, "s_cur")
        dev = s_cur_c.device
        B, N, _ = s_cur_c.shape
        attr_c, dens_c = ops._f32(attr.detach(), dev), ops._f32(dens.detach(), dev)
        wpack = owner.packed_weights(dev)
        lib = _lib.load()
        nbytes = lib.pile_train_tape_bytes(B, N)
        if nbytes < 0:
            raise _lib.PileLibraryError("unsupported sizes B=%d N=%d" % (B, N))
        tape = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        pn = None
        if particle_nums is not None:
            pn = torch.as_tensor(particle_nums).to(device=dev, dtype=torch.int32).contiguous()
        if rel is None:
            out = ops.train_forward_raw(wpack, attr_c, dens_c, s_cur_c, s_delta_c, owner.adj_thresh, pn, tape)
        else:
            out = ops.train_forward_relations_raw(wpack, attr_c, dens_c, s_cur_c, s_delta_c, rel, tape)
        owner.last_relations_buffer = (tape, 2, B, N)
        ctx.save_for_backward(wpack, dens_c, tape)
        ctx.dims, ctx.shapes = (B, N), [tuple(p.shape) for p in params]
        return out

    @staticmethod
    def backward(ctx, g):
        wpack, dens, tape = ctx.saved_tensors
        B, N = ctx.dims
        g = ops._f32(g)
        lib = _lib.load()
        total = lib.pile_train_grad_offset(len(ctx.shapes))
        grads = torch.zeros(total, dtype=torch.float32, device=g.device)
        scratch = torch.empty(lib.pile_train_scratch_bytes(B, N), dtype=torch.uint8, device=g.device)
        g_s, g_sd = ops.train_backward_raw(wpack, dens, tape, B, N, g, grads, scratch)
        out = []
        for i, shape in enumerate(ctx.shapes):
            off, end = lib.pile_train_grad_offset(i), lib.pile_train_grad_offset(i + 1)
            out.append(grads[off:end].view(shape))
        return (g_s, g_sd, None, None, None, None, None) + tuple(out)


class _GeneralStepFn(torch.autograd.Function):
    """One model step on the general-width engine (any nf_effect <= 256): input gradients always, weight gradients
    when a parameter requires grad."""

    @staticmethod
    def forward(ctx, s_cur, s_delta, attr, dens, owner, particle_nums, rel, *params):
        s_cur_c, s_delta_c = ops._f32(s_cur.detach()), ops._f32(s_delta.detach())
        ops._require_cuda(s_cur_c, "s_cur")
        dev = s_cur_c.device
        B, N, _ = s_cur_c.shape
        H = owner.nf_effect
        attr_c, dens_c = ops._f32(attr.detach(), dev), ops._f32(dens.detach(), dev)
        wpack = owner.packed_weights(dev)
        nbytes = _lib.load().pile_general_tape_bytes(B, N, H)
        if nbytes < 0:
            raise _lib.PileLibraryError("unsupported sizes B=%d N=%d nf_effect=%d" % (B, N, H))
        tape = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        pn = None
        if particle_nums is not None:
            pn = torch.as_tensor(particle_nums).to(device=dev, dtype=torch.int32).contiguous()
        # nothing asks for a gradient (torch.no_grad(): MPPI rollouts, plain predictions): the hoisted forward, whose tape
        # keeps the relation lists but not what a backward pass would read (decided by general_step: inside forward the
        # grad mode is already off and needs_input_grad ignores it)
        infer = rel is None and bool(getattr(owner, "_general_infer", False))
        out = ops.general_forward_raw(wpack, H, attr_c, dens_c, s_cur_c, s_delta_c, owner.adj_thresh, pn, tape, rel,
                                      inference=infer)
        owner.last_relations_buffer = (tape, ("general", H), B, N)
        ctx.save_for_backward(wpack, dens_c, tape)
        ctx.dims, ctx.shapes = (B, N, H), [tuple(p.shape) for p in params]
        ctx.wgrad = any(p.requires_grad for p in params)
        return out

    @staticmethod
    def backward(ctx, g):
        wpack, dens, tape = ctx.saved_tensors
        B, N, H = ctx.dims
        g = ops._f32(g)
        lib = _lib.load()
        grads = None
        if ctx.wgrad:
            grads = torch.zeros(lib.pile_general_grad_offset(len(ctx.shapes), H), dtype=torch.float32, device=g.device)
        scratch = torch.empty(lib.pile_general_scratch_bytes(B, N, H), dtype=torch.uint8, device=g.device)
        g_s, g_sd = ops.general_backward_raw(wpack, H, dens, tape, B, N, g, grads, scratch)
        out = [None] * len(ctx.shapes)
        if grads is not None:
            for i, shape in enumerate(ctx.shapes):
                off, end = lib.pile_general_grad_offset(i, H), lib.pile_general_grad_offset(i + 1, H)
                out[i] = grads[off:end].view(shape)
        return (g_s, g_sd, None, None, None, None, None) + tuple(out)


def general_step(s_cur, s_delta, attr, dens, owner, particle_nums, rel, *params):
    """_GeneralStepFn with the inference decision made where the grad mode is still visible."""
    wants = torch.is_grad_enabled() and any(isinstance(x, torch.Tensor) and x.requires_grad
                                            for x in (s_cur, s_delta, attr, dens) + tuple(params))
    owner._general_infer = not wants
    try:
        return _GeneralStepFn.apply(s_cur, s_delta, attr, dens, owner, particle_nums, rel, *params)
    finally:
        owner._general_infer = False


class PropModuleDiffDen(nn.Module):
    """Propagation network (reference model/gnn_dyn.py:113-198), weights only + CUDA forward."""

    def __init__(self, config, use_gpu=False):
        super().__init__()
        self.config = config
        nf = config['train']['particle']['nf_effect']
        self.nf_effect = nf
        # 64 (config/mpc/config.yaml:91) runs on the planner engines (FP32 tiles / tcgen05); any other width on the
        # general-width engine (csrc/general.cu)
        self.planner_engines = nf == 64
        if not 1 <= nf <= ops.GENERAL_MAX_WIDTH:
            raise _lib.PileLibraryError("nf_effect=%d: supported widths are 1..%d" % (nf, ops.GENERAL_MAX_WIDTH))
        self.add_delta = config['train']['particle']['add_delta']
        self.use_gpu = use_gpu
        # construction order = reference order, so torch.manual_seed(s) gives identical initial weights
        self.particle_encoder = _Stack("model", [(5, nf), (nf, nf)])
        self.relation_encoder = _Stack("model", [(6, nf), (nf, nf), (nf, nf)])
        self.particle_propagator = _Named(linear=(2 * nf + 1, nf))
        self.relation_propagator = _Named(linear=(3 * nf + 1, nf))
        self.particle_predictor = _Named(linear_0=(nf, nf), linear_1=(nf, 3))
        self.workspace = ops.Workspace()
        self.adj_thresh = config['train']['particle']['adj_thresh']
        self.last_relations_buffer = None
        self._packed = {}

    # ---- weights -> packed device buffer, re-packed only when a parameter changed --------------
    def packed_weights(self, device):
        params = list(self.named_parameters())
        stamp = (self.planner_engines,) + tuple((p.data_ptr(), p._version) for _, p in params)
        hit = self._packed.get(str(device))
        if hit is None or hit[0] != stamp:
            state = {"model." + k: p for k, p in params}
            packed = (ops.pack_weights(state, device) if self.planner_engines
                      else ops.pack_weights_general(state, device, self.nf_effect))
            self._packed[str(device)] = (stamp, packed)
        return self._packed[str(device)][1]

    def forward(self, a_cur, s_cur, s_delta, Rr, Rs, particle_dens, verbose=False):
        rel = Rr if isinstance(Rr, ops.Relations) else ops.Relations.from_dense(Rr, Rs)
        params = [p for _, p in self.named_parameters()]
        if not self.planner_engines:
            return general_step(s_cur, s_delta, a_cur, particle_dens, self, None, rel, *params)
        wants_wgrad = torch.is_grad_enabled() and any(p.requires_grad for p in params)
        if wants_wgrad or rel.max_degree > ops.KMAX:
            # training (weight gradients), or relation lists denser than the planner's engines take: general kernels
            return _TrainStepFn.apply(s_cur, s_delta, a_cur, particle_dens, self, None, rel, *params)
        return _StepFn.apply(s_cur, s_delta, a_cur, particle_dens, self, rel, None)


class PropNetDiffDenModel(nn.Module):
    """reference model/gnn_dyn.py:200-254."""

    def __init__(self, config, use_gpu=False):
        super().__init__()
        self.config = config
        self.adj_thresh = config['train']['particle']['adj_thresh']
        self.model = PropModuleDiffDen(config, use_gpu)

    def predict_one_step(self, a_cur, s_cur, s_delta, particle_dens, particle_nums=None):
        assert type(a_cur) == torch.Tensor
        assert type(s_cur) == torch.Tensor
        assert type(s_delta) == torch.Tensor
        assert a_cur.shape == s_cur.shape[:2]
        assert s_cur.shape == s_delta.shape
        self.model.adj_thresh = self.adj_thresh
        params = [p for _, p in self.model.named_parameters()]
        if not self.model.planner_engines:
            return general_step(s_cur, s_delta, a_cur, particle_dens, self.model, particle_nums, None, *params)
        if torch.is_grad_enabled() and any(p.requires_grad for p in params):
            # training: weight gradients wanted (train/train_gnn_dyn.py:150-199)
            return _TrainStepFn.apply(s_cur, s_delta, a_cur, particle_dens, self.model, particle_nums, None, *params)
        return _StepFn.apply(s_cur, s_delta, a_cur, particle_dens, self.model, None, particle_nums)

    def relations_of_last_step(self):
        """Relation lists built by the most recent predict_one_step (for parity checks)."""
        buf, is_tape, B, N = self.model.last_relations_buffer
        return ops.relations_from_buffer(buf, is_tape, B, N)
