"""Thin torch-facing wrappers over the libpilegnn C ABI (include/pile_gnn.h).

torch is used for device memory and streams only; every compute call goes to the CUDA
library and raises if it is missing or fails -- there is no CPU or eager fallback.
"""
import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib

KMAX = 10

# must list the slots of enum WSlot (csrc/common.cuh) in order; checked against the library at pack time
WSLOTS = ["W_PE0T", "B_PE0", "W_PE1T", "B_PE1", "W_RE0T", "B_RE0", "W_RE1T", "B_RE1", "W_RE2T", "B_RE2",
          "W_ET", "W_RT", "W_ST", "WD_RP", "B_RP", "W_PT", "W_AT", "WD_PP", "B_PP", "W_V0T", "B_V0", "W_V1T", "B_V1",
          "W_PE0", "W_PE1", "W_RE0", "W_RE1", "W_RE2", "W_E", "W_R", "W_S", "W_P", "W_A", "W_V0", "W_V1",
          "TC_EDGE", "TC_NODE", "TC_EDGE2", "TC_BWD_EDGE", "TC_BWD_NODE"]

CKPT_KEYS = [  # reference checkpoint layout, SURVEY.md §8b
    "model.particle_encoder.model.0", "model.particle_encoder.model.2",
    "model.relation_encoder.model.0", "model.relation_encoder.model.2", "model.relation_encoder.model.4",
    "model.particle_propagator.linear", "model.relation_propagator.linear",
    "model.particle_predictor.linear_0", "model.particle_predictor.linear_1"]


def set_tensor_cores(mode):
    """Select the GEMM engine: 0/False = FP32 CUDA cores, 1/True = tcgen05 tiles, 2 = tcgen05 tiles with the
    relation encoder's activations in tensor memory.  Returns the previous setting."""
    return int(_lib.load().pile_set_tensor_cores(int(mode)))


def get_tensor_cores():
    """The GEMM engine currently selected (see set_tensor_cores)."""
    return int(_lib.load().pile_get_tensor_cores())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _f32(t, device=None):
    if not isinstance(t, torch.Tensor):
        t = torch.as_tensor(t)
    if device is not None and t.device != device:
        t = t.to(device)
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _require_cuda(t, name):
    if not t.is_cuda:
        raise _lib.PileLibraryError("%s must live on a CUDA device: the pile-GNN path has no CPU implementation" % name)


def _pad_rows(m, rows):
    out = m.new_zeros(rows, m.shape[1])
    out[:m.shape[0]] = m
    return out


def _pad_cols(m, cols):
    out = m.new_zeros(m.shape[0], cols)
    out[:, :m.shape[1]] = m
    return out


def tc_operand(w):
    """[n_out][n_in] fp32 matrix -> bf16 (hi | lo) images in the canonical K-major UMMA layout (csrc/tc.cuh):
    element (n, k) at bf16 index (k//8)*(8*n_out) + (n//8)*64 + (n%8)*8 + (k%8); returned as float32 words."""
    n_out, n_in = w.shape
    hi = w.to(torch.bfloat16)
    lo = (w - hi.float()).to(torch.bfloat16)

    def canon(x):
        return x.view(n_out // 8, 8, n_in // 8, 8).permute(2, 0, 1, 3).contiguous().reshape(-1)
    return torch.cat([canon(hi), canon(lo)]).view(torch.float32)


def tc_augmented(w, k_total, extra_cols):
    """[n_out][k] weight -> [n_out][k_total] with `extra_cols` (bias, density column, ...) appended at k, zeros after."""
    out = w.new_zeros(w.shape[0], k_total)
    out[:, :w.shape[1]] = w
    for i, c in enumerate(extra_cols):
        out[:, w.shape[1] + i] = c
    return out


def pack_weights(state, device):
    """18 checkpoint tensors -> the packed float buffer the kernels read (layout: csrc/common.cuh)."""
    lib = _lib.load()
    H = lib.pile_nf_effect()
    g = {k: _f32(v.detach(), device) for k, v in state.items()}

    def w(name):
        return g[name + ".weight"], g[name + ".bias"]
    pe0, bpe0 = w(CKPT_KEYS[0]); pe1, bpe1 = w(CKPT_KEYS[1])
    re0, bre0 = w(CKPT_KEYS[2]); re1, bre1 = w(CKPT_KEYS[3]); re2, bre2 = w(CKPT_KEYS[4])
    pp, bpp = w(CKPT_KEYS[5]); rp, brp = w(CKPT_KEYS[6])
    v0, bv0 = w(CKPT_KEYS[7]); v1, bv1 = w(CKPT_KEYS[8])
    if pe1.shape != (H, H) or rp.shape != (H, 3 * H + 1) or pp.shape != (H, 2 * H + 1):
        raise _lib.PileLibraryError("libpilegnn is compiled for nf_effect=%d, checkpoint has %d" % (H, pe1.shape[0]))
    blocks = {
        "W_PE0T": _pad_rows(pe0.t(), 8), "B_PE0": bpe0, "W_PE1T": pe1.t(), "B_PE1": bpe1,
        "W_RE0T": _pad_rows(re0.t(), 8), "B_RE0": bre0, "W_RE1T": re1.t(), "B_RE1": bre1, "W_RE2T": re2.t(), "B_RE2": bre2,
        "W_ET": rp[:, 0:H].t(), "W_RT": rp[:, H:2 * H].t(), "W_ST": rp[:, 2 * H:3 * H].t(), "WD_RP": rp[:, 3 * H], "B_RP": brp,
        "W_PT": pp[:, 0:H].t(), "W_AT": pp[:, H:2 * H].t(), "WD_PP": pp[:, 2 * H], "B_PP": bpp,
        "W_V0T": v0.t(), "B_V0": bv0, "W_V1T": _pad_cols(v1.t(), 4), "B_V1": torch.cat([bv1, bv1.new_zeros(1)]),
        "W_PE0": _pad_cols(pe0, 8), "W_PE1": pe1, "W_RE0": _pad_cols(re0, 8), "W_RE1": re1, "W_RE2": re2,
        "W_E": rp[:, 0:H], "W_R": rp[:, H:2 * H], "W_S": rp[:, 2 * H:3 * H], "W_P": pp[:, 0:H], "W_A": pp[:, H:2 * H],
        "W_V0": v0, "W_V1": _pad_rows(v1, 4),
        # relation encoder on tcgen05: biases (and the density column of the last layer) ride in K chunk 8,
        # which the kernel multiplies with the constant activation chunk (1, d, 0, ...)
        "TC_EDGE": torch.cat([tc_operand(tc_augmented(re0, 16, [bre0])),
                              tc_operand(tc_augmented(re1, 80, [bre1])),
                              tc_operand(tc_augmented(re2, 80, [bre2])),
                              tc_operand(tc_augmented(rp[:, 0:H], 80, [brp, rp[:, 3 * H]]))]),
        # particle kernels on tcgen05 (layout: csrc/node_tc.cu)
        "TC_NODE": torch.cat([tc_operand(tc_augmented(pe0, 16, [bpe0])),
                              tc_operand(tc_augmented(pe1, 80, [bpe1])),
                              tc_operand(tc_augmented(pp[:, 0:H], 80, [bpp, pp[:, 2 * H]])),
                              tc_operand(torch.cat([rp[:, H:2 * H], rp[:, 2 * H:3 * H]], dim=0).contiguous()),
                              tc_operand(pp[:, H:2 * H].contiguous()),
                              tc_operand(tc_augmented(v0, 80, [bv0])),
                              tc_operand(tc_augmented(_pad_rows(v1, 16), 80, [torch.cat([bv1, bv1.new_zeros(13)])]))]),
        # relation encoder with the activation operand in tensor memory (csrc/edge_tmem.cu): plain weights,
        # biases are pre-loaded into the accumulator
        "TC_EDGE2": torch.cat([tc_operand(tc_augmented(re0, 16, [])), tc_operand(re1), tc_operand(re2),
                               tc_operand(rp[:, 0:H].contiguous())]),
        # relation-encoder dgrad on tcgen05 (csrc/bwd_tc.cu): B operand = W^T
        "TC_BWD_EDGE": torch.cat([tc_operand(rp[:, 0:H].t().contiguous()), tc_operand(re2.t().contiguous()),
                                  tc_operand(re1.t().contiguous()),
                                  tc_operand(_pad_rows(re0[:, 2:5].t().contiguous(), 16))]),
        # particle-side dgrad on tcgen05 (csrc/bwd_node_tc.cu): W_a^T, W_r^T, W_s^T, W_p^T, PE1^T, PE0[:, 0:3]^T, V0^T, V1^T
        "TC_BWD_NODE": torch.cat([tc_operand(pp[:, H:2 * H].t().contiguous()), tc_operand(rp[:, H:2 * H].t().contiguous()),
                                  tc_operand(rp[:, 2 * H:3 * H].t().contiguous()), tc_operand(pp[:, 0:H].t().contiguous()),
                                  tc_operand(pe1.t().contiguous()),
                                  tc_operand(_pad_rows(pe0[:, 0:3].t().contiguous(), 16)),
                                  tc_operand(v0.t().contiguous()), tc_operand(_pad_cols(v1.t().contiguous(), 16))]),
    }
    if lib.pile_wpack_num_slots() != len(WSLOTS):
        raise _lib.PileLibraryError("weight-slot table out of sync with libpilegnn")
    out = torch.zeros(lib.pile_wpack_total(), dtype=torch.float32, device=device)
    for i, name in enumerate(WSLOTS):
        blk = blocks[name].contiguous().reshape(-1)
        off, size = lib.pile_wpack_slot_offset(i), lib.pile_wpack_slot_size(i)
        if blk.numel() != size:
            raise _lib.PileLibraryError("slot %s: %d floats, library expects %d" % (name, blk.numel(), size))
        out[off:off + size] = blk
    return out


GENERAL_SLOTS = WSLOTS[:35]          # the non-tensor-core slots: what the general-width engine reads
GENERAL_MAX_WIDTH = 256


def pack_weights_general(state, device, H):
    """18 checkpoint tensors of ANY hidden width H <= 256 -> the packed buffer of the general-width engine
    (csrc/general.cu): the same forward / backward weight images as `pack_weights`, zero-padded to Hp = 64 * ceil(H / 64)
    so that padded channels stay exactly 0 through every layer."""
    lib = _lib.load()
    if not 1 <= H <= GENERAL_MAX_WIDTH:
        raise _lib.PileLibraryError("nf_effect=%d: the general-width engine takes 1..%d" % (H, GENERAL_MAX_WIDTH))
    Hp = (H + 63) // 64 * 64
    g = {k: _f32(v.detach(), device) for k, v in state.items()}

    def w(name):
        return g[name + ".weight"], g[name + ".bias"]
    pe0, bpe0 = w(CKPT_KEYS[0]); pe1, bpe1 = w(CKPT_KEYS[1])
    re0, bre0 = w(CKPT_KEYS[2]); re1, bre1 = w(CKPT_KEYS[3]); re2, bre2 = w(CKPT_KEYS[4])
    pp, bpp = w(CKPT_KEYS[5]); rp, brp = w(CKPT_KEYS[6])
    v0, bv0 = w(CKPT_KEYS[7]); v1, bv1 = w(CKPT_KEYS[8])
    if pe1.shape != (H, H) or rp.shape != (H, 3 * H + 1) or pp.shape != (H, 2 * H + 1):
        raise _lib.PileLibraryError("checkpoint does not have nf_effect=%d" % H)

    def sq(m):          # [r][c] -> zero-padded [Hp][Hp]
        return _pad_cols(_pad_rows(m, Hp), Hp)

    def vec(v):
        out = v.new_zeros(Hp)
        out[:v.shape[0]] = v
        return out
    blocks = {
        "W_PE0T": _pad_cols(_pad_rows(pe0.t(), 8), Hp), "B_PE0": vec(bpe0), "W_PE1T": sq(pe1.t()), "B_PE1": vec(bpe1),
        "W_RE0T": _pad_cols(_pad_rows(re0.t(), 8), Hp), "B_RE0": vec(bre0), "W_RE1T": sq(re1.t()), "B_RE1": vec(bre1),
        "W_RE2T": sq(re2.t()), "B_RE2": vec(bre2),
        "W_ET": sq(rp[:, 0:H].t()), "W_RT": sq(rp[:, H:2 * H].t()), "W_ST": sq(rp[:, 2 * H:3 * H].t()),
        "WD_RP": vec(rp[:, 3 * H]), "B_RP": vec(brp),
        "W_PT": sq(pp[:, 0:H].t()), "W_AT": sq(pp[:, H:2 * H].t()), "WD_PP": vec(pp[:, 2 * H]), "B_PP": vec(bpp),
        "W_V0T": sq(v0.t()), "B_V0": vec(bv0), "W_V1T": _pad_cols(_pad_rows(v1.t(), Hp), 4),
        "B_V1": torch.cat([bv1, bv1.new_zeros(1)]),
        "W_PE0": _pad_cols(_pad_rows(pe0, Hp), 8), "W_PE1": sq(pe1), "W_RE0": _pad_cols(_pad_rows(re0, Hp), 8),
        "W_RE1": sq(re1), "W_RE2": sq(re2),
        "W_E": sq(rp[:, 0:H]), "W_R": sq(rp[:, H:2 * H]), "W_S": sq(rp[:, 2 * H:3 * H]), "W_P": sq(pp[:, 0:H]),
        "W_A": sq(pp[:, H:2 * H]), "W_V0": sq(v0), "W_V1": _pad_cols(_pad_rows(v1, 4), Hp),
    }
    total = lib.pile_general_wpack_slot_offset(len(GENERAL_SLOTS), H)
    out = torch.zeros(total, dtype=torch.float32, device=device)
    for i, name in enumerate(GENERAL_SLOTS):
        blk = blocks[name].contiguous().reshape(-1)
        off, end = lib.pile_general_wpack_slot_offset(i, H), lib.pile_general_wpack_slot_offset(i + 1, H)
        if blk.numel() != end - off:
            raise _lib.PileLibraryError("slot %s: %d floats, library expects %d" % (name, blk.numel(), end - off))
        out[off:end] = blk
    return out


def cam_matrix12(cam_extrinsic):
    """Rows 0..2 of the world->camera(OpenCV) matrix the reference rebuilds per call (planners.py:197-203)."""
    flip = np.diag([1.0, -1.0, -1.0, 1.0])
    m = np.linalg.inv(np.matmul(np.linalg.inv(np.asarray(cam_extrinsic, dtype=np.float64)), flip))
    return [float(np.float32(v)) for v in m[:3].reshape(-1)]


class Pusher:
    """Frame of the push handed to the pusher-model kernels (`pile_pusher`, include/pile_gnn.h).

    `Pusher.sim(cam_extrinsic, global_scale)`: the simulator's camera frame (planners.py:192-257);
    `Pusher.real(s2r_scale, wkspc_center_x, wkspc_center_y)`: the real robot (gen_s_delta_irl, planners.py:259-300)."""

    def __init__(self, kind, cam12=None, global_scale=1.0, s2r_scale=1.0, center=(0.0, 0.0)):
        self.kind = int(kind)
        self.cam12 = [0.0] * 12 if cam12 is None else [float(v) for v in cam12]
        self.global_scale = float(global_scale)
        self.s2r_scale = float(s2r_scale)
        self.center = (float(center[0]), float(center[1]))
        st = _lib.PusherStruct()
        st.kind = self.kind
        for i, v in enumerate(self.cam12):
            st.cam_m12[i] = v
        st.global_scale, st.s2r_scale = self.global_scale, self.s2r_scale
        st.wkspc_center_x, st.wkspc_center_y = self.center
        self.struct = st

    @staticmethod
    def sim(cam_extrinsic, global_scale):
        return Pusher(0, cam_matrix12(cam_extrinsic), global_scale)

    @staticmethod
    def real(s2r_scale, wkspc_center_x, wkspc_center_y):
        return Pusher(1, None, 1.0, s2r_scale, (wkspc_center_x, wkspc_center_y))

    def ref(self):
        return C.byref(self.struct)

    def signature(self):
        return (self.kind, tuple(self.cam12), self.global_scale, self.s2r_scale, self.center)


@dataclass
class Relations:
    """Compact relation lists of a batch: the Rr/Rs-equivalent (SURVEY.md §8b 'Forward')."""
    rowptr: torch.Tensor   # [B, N+1] int32, offsets local to the sample
    col: torch.Tensor      # [B, 10N] int32 sender of relation e
    row: torch.Tensor      # [B, 10N] int32 receiver of relation e
    max_degree: int = KMAX  # most relations any particle receives (the planner's engines handle up to KMAX)

    @property
    def n_rel(self):
        return self.rowptr[:, -1]

    def edge_sets(self):
        """list over samples of int arrays [E_b, 2] = (receiver, sender), in storage order."""
        rp, col, row = self.rowptr.cpu().numpy(), self.col.cpu().numpy(), self.row.cpu().numpy()
        return [np.stack([row[b, :rp[b, -1]], col[b, :rp[b, -1]]], axis=1) for b in range(rp.shape[0])]

    def to_dense(self, dtype=torch.float32):
        """One-hot Rr, Rs [B, n_rel, N] laid out like the reference (gnn_dyn.py:242-251)."""
        B, N = self.rowptr.shape[0], self.rowptr.shape[1] - 1
        n_rel = int(self.n_rel.max())
        Rr = torch.zeros(B, n_rel, N, dtype=dtype, device=self.col.device)
        Rs = torch.zeros_like(Rr)
        slot = torch.arange(n_rel, device=self.col.device)[None].expand(B, n_rel)
        live = slot < self.n_rel[:, None]
        b_idx = torch.arange(B, device=self.col.device)[:, None].expand(B, n_rel)
        Rr[b_idx[live], slot[live], self.row[:, :n_rel][live].long()] = 1
        Rs[b_idx[live], slot[live], self.col[:, :n_rel][live].long()] = 1
        return Rr, Rs

    @staticmethod
    def from_dense(Rr, Rs):
        """One-hot [B, rel, N] (any relation order, zero rows = padding) -> receiver-grouped lists."""
        B, n_rel, N = Rr.shape
        dev = Rr.device
        live = Rr.sum(2) > 0
        recv = Rr.argmax(2)
        send = Rs.argmax(2)
        key = torch.where(live, recv * N + send, torch.full_like(recv, N * N))
        order = key.argsort(dim=1, stable=True)
        recv, send, live = recv.gather(1, order), send.gather(1, order), live.gather(1, order)
        if n_rel > KMAX * N:
            raise ValueError("more than %d relations per sample" % (KMAX * N))
        col = torch.zeros(B, KMAX * N, dtype=torch.int32, device=dev)
        row = torch.zeros(B, KMAX * N, dtype=torch.int32, device=dev)
        col[:, :n_rel] = torch.where(live, send, torch.zeros_like(send)).int()
        row[:, :n_rel] = torch.where(live, recv, torch.zeros_like(recv)).int()
        counts = torch.zeros(B, N + 1, dtype=torch.int64, device=dev)
        counts.scatter_add_(1, torch.where(live, recv + 1, torch.zeros_like(recv)), live.long())
        counts[:, 0] = 0
        # the planner's engines keep one receiver's relations in a KMAX-row shared-memory slab (the reference's own
        # builder never emits more: topk(k=10), gnn_dyn.py:231); lists with higher degrees are routed to the general
        # kernels of the training path by PropModuleDiffDen.forward
        max_degree = int(counts.max()) if n_rel > 0 else 0
        return Relations(counts.cumsum(1).int().contiguous(), col, row, max(max_degree, 1))


class Workspace:
    """Device scratch reused across calls (the library never allocates).  The dynamic-resolution MPC loop changes
    the particle count every step, so only the few most recently used (B, N) sizes are kept: anything older is
    dropped (a holder such as RolloutEngine or a captured planner loop keeps its own reference alive)."""
    KEEP = 3

    def __init__(self):
        self._scratch = {}
        self._bwd = {}

    def _get(self, cache, key, nbytes_fn):
        hit = cache.pop(key, None)
        if hit is None:
            n = nbytes_fn(key[0], key[1])
            if n < 0:
                raise _lib.PileLibraryError("unsupported sizes B=%d N=%d" % (key[0], key[1]))
            while len(cache) >= self.KEEP:
                cache.pop(next(iter(cache)))            # least recently used first (dict keeps insertion order)
            hit = torch.empty(n, dtype=torch.uint8, device=key[2])
        cache[key] = hit
        return hit

    def scratch(self, B, N, device):
        return self._get(self._scratch, (int(B), int(N), str(device)), _lib.load().pile_step_scratch_bytes)

    def bwd(self, B, N, device):
        return self._get(self._bwd, (int(B), int(N), str(device)), _lib.load().pile_bwd_scratch_bytes)


def new_tape(B, N, T, device):
    n = _lib.load().pile_tape_step_bytes(B, N)
    if n < 0:
        raise _lib.PileLibraryError("unsupported sizes B=%d N=%d" % (B, N))
    return torch.empty(n * T, dtype=torch.uint8, device=device)


def gen_s_delta_raw(s_cur, action, pusher):
    _require_cuda(s_cur, "s_cur")
    B, N, _ = s_cur.shape
    out = torch.empty_like(s_cur)
    _lib.check(_lib.load().pile_gen_s_delta(_lib.ptr(s_cur), _lib.ptr(action), action.stride(0), pusher.ref(), B, N,
                                            _lib.ptr(out), _stream()), "pile_gen_s_delta")
    return out


def gen_s_delta_backward_raw(s_cur, action, pusher, g_sd):
    B, N, _ = s_cur.shape
    g_s = torch.zeros_like(s_cur)
    g_a = torch.empty(B, 4, dtype=torch.float32, device=s_cur.device)
    _lib.check(_lib.load().pile_gen_s_delta_backward(
        _lib.ptr(s_cur), _lib.ptr(action), action.stride(0), pusher.ref(), B, N, _lib.ptr(g_sd), _lib.ptr(g_s),
        _lib.ptr(g_a), 4, _stream()), "pile_gen_s_delta_backward")
    return g_s, g_a


class _GenSDelta(torch.autograd.Function):
    @staticmethod
    def forward(ctx, s_cur, action, pusher):
        s_cur, action = _f32(s_cur), _f32(action)
        ctx.save_for_backward(s_cur, action)
        ctx.pusher = pusher
        return gen_s_delta_raw(s_cur, action, pusher)

    @staticmethod
    def backward(ctx, g):
        s_cur, action = ctx.saved_tensors
        g_s, g_a = gen_s_delta_backward_raw(s_cur, action, ctx.pusher, _f32(g))
        return g_s, g_a, None


def gen_s_delta(s_cur, action, pusher):
    """Differentiable pusher model (planners.py:211-257; :259-300 for a real-robot `pusher`)."""
    return _GenSDelta.apply(s_cur, action, pusher)


def build_relations(s_cur, s_delta, adj_thresh, particle_nums=None, with_transpose=False):
    """model/gnn_dyn.py:221-251 -> Relations (bit-exact relation set, torch.nonzero order).
    with_transpose: also return (trowptr [B,N+1], trecv [B,10N], tedge [B,10N]), the sender-major transpose the
    backward uses (for sender j: its receivers in ascending order and the ids of those relations)."""
    s_cur, s_delta = _f32(s_cur.detach()), _f32(s_delta.detach())
    _require_cuda(s_cur, "s_cur")
    B, N, _ = s_cur.shape
    dev = s_cur.device
    rowptr = torch.empty(B, N + 1, dtype=torch.int32, device=dev)
    col = torch.zeros(B, KMAX * N, dtype=torch.int32, device=dev)
    row = torch.zeros(B, KMAX * N, dtype=torch.int32, device=dev)
    pn = None if particle_nums is None else torch.as_tensor(particle_nums).to(device=dev, dtype=torch.int32).contiguous()
    trowptr = trecv = tedge = None
    if with_transpose:
        trowptr = torch.empty(B, N + 1, dtype=torch.int32, device=dev)
        trecv = torch.zeros(B, KMAX * N, dtype=torch.int32, device=dev)
        tedge = torch.zeros(B, KMAX * N, dtype=torch.int32, device=dev)
    _lib.check(_lib.load().pile_build_relations(_lib.ptr(s_cur), _lib.ptr(s_delta), _lib.ptr(pn), B, N, float(adj_thresh),
                                                _lib.ptr(rowptr), _lib.ptr(col), _lib.ptr(row), _lib.ptr(trowptr),
                                                _lib.ptr(trecv), _lib.ptr(tedge), _stream()), "pile_build_relations")
    rel = Relations(rowptr, col, row)
    return (rel, (trowptr, trecv, tedge)) if with_transpose else rel


def relations_from_buffer(buf, is_tape, B, N):
    """View the relation lists a step left in its scratch / tape / training-tape (is_tape == 2) buffer (no copy)."""
    lib = _lib.load()
    ps = [C.c_void_p() for _ in range(3)]
    if isinstance(is_tape, tuple):          # ("general", nf_effect): tape of the general-width engine
        _lib.check(lib.pile_general_relations_view(_lib.ptr(buf), B, N, int(is_tape[1]), *[C.byref(p) for p in ps]),
                   "pile_general_relations_view")
    elif int(is_tape) == 2:
        _lib.check(lib.pile_train_relations_view(_lib.ptr(buf), B, N, *[C.byref(p) for p in ps]), "pile_train_relations_view")
    else:
        _lib.check(lib.pile_relations_view(_lib.ptr(buf), int(is_tape), B, N, *[C.byref(p) for p in ps]), "pile_relations_view")
    base = buf.data_ptr()
    out = []
    for p, n in zip(ps, [B * (N + 1), B * KMAX * N, B * KMAX * N]):
        off = p.value - base
        out.append(buf[off:off + 4 * n].view(torch.int32))
    return Relations(out[0].view(B, N + 1), out[1].view(B, KMAX * N), out[2].view(B, KMAX * N))


def predict_step_raw(wpack, attr, dens, s_cur, s_delta, adj_thresh, particle_nums, scratch, tape):
    B, N, _ = s_cur.shape
    out = torch.empty_like(s_cur)
    _lib.check(_lib.load().pile_predict_step(_lib.ptr(wpack), _lib.ptr(attr), _lib.ptr(dens), _lib.ptr(particle_nums),
                                             _lib.ptr(s_cur), _lib.ptr(s_delta), float(adj_thresh), B, N,
                                             _lib.ptr(scratch), _lib.ptr(tape), _lib.ptr(out), _stream()),
               "pile_predict_step")
    return out


def forward_relations_raw(wpack, attr, dens, s_cur, s_delta, rel, scratch, tape):
    B, N, _ = s_cur.shape
    out = torch.empty_like(s_cur)
    _lib.check(_lib.load().pile_forward_relations(
        _lib.ptr(wpack), _lib.ptr(attr), _lib.ptr(dens), _lib.ptr(s_cur), _lib.ptr(s_delta), _lib.ptr(rel.rowptr),
        _lib.ptr(rel.col), _lib.ptr(rel.row), B, N, _lib.ptr(scratch), _lib.ptr(tape), _lib.ptr(out), _stream()),
        "pile_forward_relations")
    return out


def step_backward_raw(wpack, dens, tape, B, N, g_pred, bwd_scratch):
    g_s = torch.empty(B, N, 3, dtype=torch.float32, device=g_pred.device)
    g_sd = torch.empty_like(g_s)
    _lib.check(_lib.load().pile_step_backward(_lib.ptr(wpack), _lib.ptr(dens), _lib.ptr(tape), B, N, _lib.ptr(g_pred),
                                              _lib.ptr(g_s), _lib.ptr(g_sd), _lib.ptr(bwd_scratch), _stream()),
               "pile_step_backward")
    return g_s, g_sd


def rollout_forward_raw(wpack, attr, dens, s0, actions, pusher, adj_thresh, scratch, tape, out=None):
    B, N, _ = s0.shape
    T = actions.shape[1]
    if out is None:
        out = torch.empty(B, T, N, 3, dtype=torch.float32, device=s0.device)
    _lib.check(_lib.load().pile_rollout_forward(
        _lib.ptr(wpack), _lib.ptr(attr), _lib.ptr(dens), _lib.ptr(s0), _lib.ptr(actions), pusher.ref(),
        float(adj_thresh), B, N, T, _lib.ptr(scratch), _lib.ptr(tape), _lib.ptr(out), _stream()),
        "pile_rollout_forward")
    return out


def rollout_backward_raw(wpack, dens, s0, actions, pusher, tape, states, g_states, bwd_scratch, out=None):
    B, T, N, _ = states.shape
    g_act = out if out is not None else torch.empty(B, T, 4, dtype=torch.float32, device=states.device)
    _lib.check(_lib.load().pile_rollout_backward(
        _lib.ptr(wpack), _lib.ptr(dens), _lib.ptr(s0), _lib.ptr(actions), pusher.ref(), B, N, T, _lib.ptr(tape), _lib.ptr(states), _lib.ptr(g_states), _lib.ptr(bwd_scratch), _lib.ptr(g_act),
        _stream()), "pile_rollout_backward")
    return g_act


def reward_raw(states, n_states, state_stride, N, goal_img, goal_coor, cam_params, offset, normalize, want_argmin=False,
               out=None, arg=None):
    M = goal_coor.shape[0]
    Hh, Ww = goal_img.shape
    if out is None:
        out = torch.empty(n_states, dtype=torch.float32, device=goal_img.device)
    if arg is None and want_argmin:
        arg = torch.empty(n_states, M, dtype=torch.int32, device=goal_img.device)
    _lib.check(_lib.load().pile_reward(_lib.ptr(states), n_states, state_stride, N, _lib.ptr(goal_img), Hh, Ww,
                                       _lib.ptr(goal_coor), M, _lib.host_floats(cam_params), float(offset[0]),
                                       float(offset[1]), int(bool(normalize)), _lib.ptr(out), _lib.ptr(arg), _stream()),
               "pile_reward")
    return out, arg


def reward_backward_raw(states, n_states, state_stride, N, goal_img, goal_coor, cam_params, offset, normalize,
                        g_reward, argmin, g_states, g_stride, accumulate):
    M = goal_coor.shape[0]
    Hh, Ww = goal_img.shape
    _lib.check(_lib.load().pile_reward_backward(
        _lib.ptr(states), n_states, state_stride, N, _lib.ptr(goal_img), Hh, Ww, _lib.ptr(goal_coor), M,
        _lib.host_floats(cam_params), float(offset[0]), float(offset[1]), int(bool(normalize)), _lib.ptr(g_reward),
        _lib.ptr(argmin), _lib.ptr(g_states), g_stride, int(bool(accumulate)), _stream()), "pile_reward_backward")


def mppi_partials(reward, acts, reward_weight):
    """reward [S], acts [S,T,4] -> one record [2+4T] = (max z, sum exp, sum exp*act) for this device."""
    lib = _lib.load()
    S, T = acts.shape[0], acts.shape[1]
    P = lib.pile_mppi_num_chunks(S)
    part = torch.empty(P, 2 + 4 * T, dtype=torch.float32, device=acts.device)
    _lib.check(lib.pile_mppi_partials(_lib.ptr(reward), _lib.ptr(acts), S, T, float(reward_weight), _lib.ptr(part),
                                      _stream()), "pile_mppi_partials")
    return mppi_combine(part, T)


def mppi_combine(parts, T):
    """parts [P, 2+4T] -> merged record [2+4T] (log-sum-exp rescale)."""
    parts = parts.contiguous()
    out = torch.empty(2 + 4 * T, dtype=torch.float32, device=parts.device)
    _lib.check(_lib.load().pile_mppi_combine(_lib.ptr(parts), parts.shape[0], T, _lib.ptr(out), _stream()),
               "pile_mppi_combine")
    return out


def fps(pts, count, init_idx=0):
    """Farthest-point sampling on the GPU (utils.fps_np contract): pts [n, dim] or [sets, n, dim] float32 CUDA
    -> (picked points [.., count, dim], indices [.., count] int32, covering radius [..])."""
    pts = _f32(pts)
    _require_cuda(pts, "pts")
    single = pts.dim() == 2
    p3 = pts[None] if single else pts
    S, n, dim = p3.shape
    dev = pts.device
    gap = torch.empty(S, n, dtype=torch.float32, device=dev)
    idx = torch.empty(S, count, dtype=torch.int32, device=dev)
    out = torch.empty(S, count, dim, dtype=torch.float32, device=dev)
    rad = torch.empty(S, dtype=torch.float32, device=dev)
    _lib.check(_lib.load().pile_fps(_lib.ptr(p3.contiguous()), S, n, dim, int(count), int(init_idx), _lib.ptr(gap),
                                    _lib.ptr(idx), _lib.ptr(out), _lib.ptr(rad), _stream()), "pile_fps")
    return (out[0], idx[0], rad[0]) if single else (out, idx, rad)


def adam_clamp(actions, grad, exp_avg, exp_avg_sq, step, lr, lo4, hi4, betas=(0.9, 0.999), eps=1e-8):
    """In-place torch.optim.Adam update of `actions` [..., 4] for 1-based `step`, then clamp to the box."""
    _lib.check(_lib.load().pile_adam_clamp(_lib.ptr(actions), _lib.ptr(grad), _lib.ptr(exp_avg), _lib.ptr(exp_avg_sq),
                                           actions.numel(), int(step), float(lr), float(betas[0]), float(betas[1]),
                                           float(eps), _lib.host_floats(lo4), _lib.host_floats(hi4), _stream()),
               "pile_adam_clamp")


def adam_clamp_dev(actions, grad, exp_avg, exp_avg_sq, iter_dev, lr, lo4, hi4, betas=(0.9, 0.999), eps=1e-8):
    """adam_clamp with the step number (= iter_dev[0] + 1) read on the device: graph-replayable."""
    _lib.check(_lib.load().pile_adam_clamp_dev(_lib.ptr(actions), _lib.ptr(grad), _lib.ptr(exp_avg),
                                               _lib.ptr(exp_avg_sq), actions.numel(), _lib.ptr(iter_dev), float(lr),
                                               float(betas[0]), float(betas[1]), float(eps), _lib.host_floats(lo4),
                                               _lib.host_floats(hi4), _stream()), "pile_adam_clamp_dev")


def counter_add(counter, delta=1):
    _lib.check(_lib.load().pile_counter_add(_lib.ptr(counter), int(delta), _stream()), "pile_counter_add")


def gd_track(reward, actions, n_sample, n_batch, T, max_reward, max_idx, best_actions, rew_mean, rew_std, iter_dev,
             stat_every=None, stat_stride=None):
    """Per-iteration best tracking + reward statistics of the GD planner on the device (planners.py:721-740);
    rew_mean / rew_std are [n_batch // stat_every, stat_stride] (one row per scene)."""
    stat_every = int(n_batch) if stat_every is None else int(stat_every)
    stat_stride = int(rew_mean.shape[-1]) if stat_stride is None else int(stat_stride)
    _lib.check(_lib.load().pile_gd_track(_lib.ptr(reward), _lib.ptr(actions), int(n_sample), int(n_batch), int(T),
                                         _lib.ptr(max_reward), _lib.ptr(max_idx), _lib.ptr(best_actions),
                                         _lib.ptr(rew_mean), _lib.ptr(rew_std), _lib.ptr(iter_dev), stat_every,
                                         stat_stride, _stream()), "pile_gd_track")


def train_forward_raw(wpack, attr, dens, s_cur, s_delta, adj_thresh, particle_nums, tape):
    """Model step that keeps every layer input in `tape` for the weight-gradient backward (csrc/train.cu)."""
    B, N, _ = s_cur.shape
    out = torch.empty_like(s_cur)
    _lib.check(_lib.load().pile_train_forward(_lib.ptr(wpack), _lib.ptr(attr), _lib.ptr(dens), _lib.ptr(particle_nums),
                                              _lib.ptr(s_cur), _lib.ptr(s_delta), float(adj_thresh), B, N, _lib.ptr(tape),
                                              _lib.ptr(out), _stream()), "pile_train_forward")
    return out


def train_forward_relations_raw(wpack, attr, dens, s_cur, s_delta, rel, tape):
    """train_forward_raw on caller-provided relation lists (any degree, at most 10 N relations per sample)."""
    B, N, _ = s_cur.shape
    out = torch.empty_like(s_cur)
    _lib.check(_lib.load().pile_train_forward_relations(
        _lib.ptr(wpack), _lib.ptr(attr), _lib.ptr(dens), _lib.ptr(s_cur), _lib.ptr(s_delta), _lib.ptr(rel.rowptr),
        _lib.ptr(rel.col), _lib.ptr(rel.row), B, N, _lib.ptr(tape), _lib.ptr(out), _stream()),
        "pile_train_forward_relations")
    return out


def train_backward_raw(wpack, dens, tape, B, N, g_pred, grads, scratch):
    """-> (g_s_cur, g_s_delta); the 18 weight gradients are accumulated into the flat buffer `grads`."""
    g_s = torch.empty(B, N, 3, dtype=torch.float32, device=g_pred.device)
    g_sd = torch.empty_like(g_s)
    _lib.check(_lib.load().pile_train_backward(_lib.ptr(wpack), _lib.ptr(dens), _lib.ptr(tape), B, N, _lib.ptr(g_pred),
                                               _lib.ptr(g_s), _lib.ptr(g_sd), _lib.ptr(grads), _lib.ptr(scratch),
                                               _stream()), "pile_train_backward")
    return g_s, g_sd


def general_forward_raw(wpack, H, attr, dens, s_cur, s_delta, adj_thresh, particle_nums, tape, rel=None, inference=False):
    """Model step of the general-width engine (csrc/general.cu, any nf_effect <= 256); relations searched or supplied.
    inference: no backward pass will follow -- the regrouped (hoisted) relation propagator, searched relations only."""
    B, N, _ = s_cur.shape
    out = torch.empty_like(s_cur)
    lib = _lib.load()
    if rel is None:
        fn, name = (lib.pile_general_forward_inference, "pile_general_forward_inference") if inference else \
            (lib.pile_general_forward, "pile_general_forward")
        _lib.check(fn(_lib.ptr(wpack), int(H), _lib.ptr(attr), _lib.ptr(dens), _lib.ptr(particle_nums),
                      _lib.ptr(s_cur), _lib.ptr(s_delta), float(adj_thresh), B, N, _lib.ptr(tape),
                      _lib.ptr(out), _stream()), name)
    else:
        _lib.check(lib.pile_general_forward_relations(
            _lib.ptr(wpack), int(H), _lib.ptr(attr), _lib.ptr(dens), _lib.ptr(s_cur), _lib.ptr(s_delta), _lib.ptr(rel.rowptr),
            _lib.ptr(rel.col), _lib.ptr(rel.row), B, N, _lib.ptr(tape), _lib.ptr(out), _stream()),
            "pile_general_forward_relations")
    return out


def general_backward_raw(wpack, H, dens, tape, B, N, g_pred, grads, scratch):
    """-> (g_s_cur, g_s_delta); grads (flat buffer of the 18 tensors) is accumulated into, None = input gradients only."""
    g_s = torch.empty(B, N, 3, dtype=torch.float32, device=g_pred.device)
    g_sd = torch.empty_like(g_s)
    _lib.check(_lib.load().pile_general_backward(_lib.ptr(wpack), int(H), _lib.ptr(dens), _lib.ptr(tape), B, N,
                                                 _lib.ptr(g_pred), _lib.ptr(g_s), _lib.ptr(g_sd), _lib.ptr(grads),
                                                 _lib.ptr(scratch), _stream()), "pile_general_backward")
    return g_s, g_sd
